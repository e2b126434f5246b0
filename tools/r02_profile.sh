set -x
# 1. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --logn 28 --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02_ncu_launches.log 2>&1
# 2. DRAM traffic of the 32 scan launches at the bench size
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:scan_(range|direct)_kernel' --csv --log-file gpurun_out/scan_traffic.csv python tools/scan_traffic.py > /dev/null 2>&1
# 3. full captures of the scan kernel at 1, 2, 3, 11, 17 bit
for w in 1 2 3 11 17; do
  ncu --set full --clock-control none --import-source on -k 'regex:scan_(range|direct)_kernel' -s 2 -c 1 -o gpurun_out/r02_ncu_scan_w$w -f python tools/ncu_family_driver.py scan $w 30 > /dev/null 2>&1
  python tools/ncu_brief.py gpurun_out/r02_ncu_scan_w$w.ncu-rep $((1<<30)) > gpurun_out/r02_ncu_scan_w$w.txt
  rm -f gpurun_out/r02_ncu_scan_w$w.ncu-rep
done
# 4. pack and pos kernels
ncu --set full --clock-control none --import-source on -k 'regex:pack_block_kernel' -s 2 -c 1 -o gpurun_out/r02_ncu_pack_w17 -f python tools/ncu_family_driver.py pack 17 28 > /dev/null 2>&1
python tools/ncu_brief.py gpurun_out/r02_ncu_pack_w17.ncu-rep $((1<<28)) > gpurun_out/r02_ncu_pack_w17.txt; rm -f gpurun_out/r02_ncu_pack_w17.ncu-rep
SEL=0.001 ncu --set full --clock-control none --import-source on -k regex:pos_expand -s 2 -c 1 -o gpurun_out/r02_ncu_pos_expand_sel001_v7 -f python tools/ncu_family_driver.py positions 17 30 > /dev/null 2>&1
python tools/ncu_brief.py gpurun_out/r02_ncu_pos_expand_sel001_v7.ncu-rep > gpurun_out/r02_ncu_pos_expand_sel001_v7.txt; rm -f gpurun_out/r02_ncu_pos_expand_sel001_v7.ncu-rep
head -3 gpurun_out/r02_ncu_scan_w1.txt
