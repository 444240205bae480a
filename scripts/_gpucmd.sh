for b in 4 8; do
echo "BATCHES=$b"; SA_GPU_COOP_BATCHES=$b SA_GPU_PROFILE=1 timeout 300 python tests/_full_probe.py 128 4 64 2>&1 | grep -oE "eig.bisect [0-9.]+|eig.inverse_iter [0-9.]+|eig.back_transform [0-9.]+|eig.large_tridiag [0-9.]+|'l[12].local_spectral': [0-9.]+|'setup': [0-9.]+" | tr '\n' ' '; echo
done
SA_GPU_ALLOC_DEBUG=1 timeout 300 python tests/_full_probe.py 128 4 64 2> gpurun_out/a.err | grep -oE "'l[12].local_spectral': [0-9.]+|'setup': [0-9.]+|pcg iters [0-9]+" | tr '\n' ' '; echo; grep alloc gpurun_out/a.err | awk '{s+=$5; n++; b+=$2} END {print n, "allocs >=64MB,", b/1024, "GB, total ms", s}'
