for v in 0 1 0 1; do
if [ $v = 1 ]; then export SA_NO_TOPOLOGY_PREFETCH=1; else unset SA_NO_TOPOLOGY_PREFETCH; fi
echo "NOPREFETCH=$v"; timeout 300 python tests/_full_probe.py 128 4 64 2>&1 | grep -oE "ml_build [0-9.]+s|'l0.local_spectral': [0-9.]+|'l0.rap': [0-9.]+|'l1.local_spectral': [0-9.]+|'l1.topology': [0-9.]+|'l2.coarse_elmats': [0-9.]+|'setup': [0-9.]+" | tr '\n' ' '; echo
done
