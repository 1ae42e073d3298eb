mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"zl_k_match_small|zl_k_parse|zl_k_enc_literals|zl_k_enc_sequences" -s 8 -c 4 -f -o gpurun_out/r02c_dictenc python tools/probe_dict.py 100000 > gpurun_out/t7_ncu.log 2>&1
ncu -i gpurun_out/r02c_dictenc.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_full_dict_compress_kernels_raw.csv 2>> gpurun_out/t7_ncu.log
ncu -i gpurun_out/r02c_dictenc.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02c_dictenc_src.csv 2>> gpurun_out/t7_ncu.log
ls -la gpurun_out/r02c_dictenc* >> gpurun_out/t7_ncu.log
rm -f gpurun_out/r02c_dictenc.ncu-rep
tail -5 gpurun_out/t7_ncu.log
