timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python scratch/c4dbg.py 320 > gpurun_out/c4dbg_synccheck.txt 2>&1
grep -v "Host Frame" gpurun_out/c4dbg_synccheck.txt | head -40
timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 python scratch/c4dbg.py 320 > gpurun_out/c4dbg_racecheck.txt 2>&1
grep -v "Host Frame" gpurun_out/c4dbg_racecheck.txt | head -80
