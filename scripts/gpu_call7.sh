source scripts/gpu_check.sh c7
./singlespmv_b200/plugin/bin/spmv_b200_crs_dev synth:lap2d5:1024 > gpurun_out/driver_c1_crs_dev.txt 2> gpurun_out/driver_c1_crs_dev.err; cat gpurun_out/driver_c1_crs_dev.txt
./singlespmv_b200/plugin/bin/spmv_b200_crs synth:lap2d5:1024 > gpurun_out/driver_c1_crs_host.txt 2>> gpurun_out/driver_c1_crs_dev.err; grep -E "Performance|Vector" gpurun_out/driver_c1_crs_host.txt
