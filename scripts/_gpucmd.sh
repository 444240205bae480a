for d in 300 300 180 180; do
echo "DIV $d"; SA_GPU_COOP_DIV=$d timeout 300 python tests/_full_probe.py 128 4 64 2>&1 | grep -oE "'l[12].local_spectral': [0-9.]+"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv
