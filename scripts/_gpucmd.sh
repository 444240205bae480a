for gb in 3.2 3.2 24 24 1.5; do
echo "CHUNK_GB $gb"; SA_GPU_CHUNK_GB=$gb timeout 300 python tests/_full_probe.py 128 4 64 2>&1 | grep -oE "'l[012].local_spectral': [0-9.]+|'setup': [0-9.]+" | tr '\n' ' '; echo
done
