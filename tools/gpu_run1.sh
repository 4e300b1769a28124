mkdir -p gpurun_out
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2f
timeout 900 ncu --set full --clock-control none -k regex:'tile_image|untile_blocks|decode_kernel|s3tc_encode|eac_encode' -f -o /tmp/misc_r2f python tools/prof_misc.py > gpurun_out/ncu_misc_r2f.log 2>&1
tail -n 2 gpurun_out/ncu_misc_r2f.log
ncu -i /tmp/misc_r2f.ncu-rep --page raw --csv > gpurun_out/misc_r2f_raw.csv
ls -la gpurun_out
