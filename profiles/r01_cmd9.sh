set -x
for m in 0 1 2; do echo "== addend mode $m"; SIVAE_TC_ADDEND=$m PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py 2>&1 | tail -4; done > gpurun_out/probe_addend.log 2>&1
cat gpurun_out/probe_addend.log
for m in 0 1; do SIVAE_TC_ADDEND=$m timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_H16_$m.log 2>&1; tail -1 gpurun_out/bench_H16_$m.log | cut -c1-200; done
