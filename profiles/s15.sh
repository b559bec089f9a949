set -u
mkdir -p gpurun_out
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_eig -s 200 -c 1 -o gpurun_out/svd_eig \
    python profiles/svd_driver.py 15 4096 1.6 decay > gpurun_out/ncu_eig.log 2>&1
ls -la gpurun_out/svd_eig.ncu-rep
