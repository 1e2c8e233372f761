timeout 300 ./tools/gemm_lab check > gpurun_out/lab_check_s19.txt 2>&1; grep -c "^ok" gpurun_out/lab_check_s19.txt; grep "bn128 split1\|FAIL\|check:" gpurun_out/lab_check_s19.txt
timeout 300 ./tools/gemm_lab time > gpurun_out/lab_time_s19.txt 2>&1; grep "split1" gpurun_out/lab_time_s19.txt | grep "N1024 K512\|N512 K1024" 
bash tools/gpu_ab.sh s19 "JB_WIDE_SPLIT=0" "JB_WIDE_SPLIT=1" "JB_WIDE_SPLIT=0" "JB_WIDE_SPLIT=1"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
