ulimit -a; free -g; nproc; grep -m1 "model name" /proc/cpuinfo
python scratch/dbg_c5.py > gpurun_out/dbg_c5_plain.txt 2>&1; echo rc=$? >> gpurun_out/dbg_c5_plain.txt
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scratch/dbg_c5.py oracle/_ref/asan/libcorto_ref.so > gpurun_out/dbg_c5_asan.txt 2>&1; echo rc=$? >> gpurun_out/dbg_c5_asan.txt
grep -v "^  File" gpurun_out/dbg_c5_plain.txt | head -20; grep -v "^  File" gpurun_out/dbg_c5_asan.txt | head -60
