python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse" 2>&1 | tail -15
python tools/probe_vi.py --n 3000000 --k 2000 --s 16 --modes rcgsp,emsp --iters 20 2>&1 | grep -E "pass-kernels|theta"
ncu --set full --clock-control none -k regex:em_sparse_pass -s 4 -c 1 -o /tmp/sp python tools/probe_vi.py --n 20000000 --k 2000 --s 16 --modes emsp --iters 3 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/sp.ncu-rep > gpurun_out/sparse_r02_summary.txt 2>&1
python tools/ncu_sass_hist.py /tmp/sp.ncu-rep 30 > gpurun_out/sparse_r02_sass.txt 2>&1
ncu -i /tmp/sp.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for i,name in enumerate(h):
        if any(t in name for t in ('l1tex__data_pipe_lsu_wavefronts','bank_conflicts','lts__t_bytes','l1tex__t_bytes','smsp__inst_executed_op_shared','dram__bytes','lts__throughput','l1tex__throughput')): print(name, r[i])
" > gpurun_out/sparse_r02_lsu.txt 2>&1
ncu --set full --clock-control none -k regex:rcg_sweep -s 8 -c 2 -o /tmp/rk python tools/probe_k.py --ks 50 --modes rcg --iters 3 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/rk.ncu-rep > gpurun_out/rcg_k50_r02_summary.txt 2>&1
python tools/ncu_sass_hist.py /tmp/rk.ncu-rep 30 > gpurun_out/rcg_k50_r02_sass.txt 2>&1
ls -la gpurun_out/
