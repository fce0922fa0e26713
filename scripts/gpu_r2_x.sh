#!/bin/bash
# round 2, call X: the bench as the driver runs it (one-time set-up before the warm-up), the whole GPU
# parity suite on the current library, decoder phases on the noisy field, ncu of the encoder's
# heaviest launches (last plane of k_lis_plane, k_lipref_emit)
mkdir -p gpurun_out /tmp/rep
echo "== bench"
python bench.py > gpurun_out/r2x_bench1.log 2>&1; tail -1 gpurun_out/r2x_bench1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d.get('warmup_extra_steps'))
print('steps', d['step_ms_each']); print('host', d['step_host_ms_each']); print('c', d['compress_ms_each'], 'd', d['decompress_ms_each'])
print('e2e', d['e2e']); print('stages', d['stages_ms']); print('roofline', d['roofline']); print(d['roofline_wavelet']); print(d['cpu_baseline']); print(d['parity'], d['gpu_launches'], d['clocks'])"
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
echo "== noisy"; timeout 300 python scripts/dec_noisy.py 512 2>&1 | grep -v "^$" | cut -c1-900 | tail -12
B="python bench.py --steps 1 --warmup 1 --settle 0 --e2e 0 --cpu-baseline 0 --check 0"
cap() {  # name, regex, skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s $skip -c 1 -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1
  python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_top_lines.py > gpurun_out/r2_ncu_${name}_source_top.txt 2>&1
  grep -E "^==|Duration|Ipc|stall (barrier|long|wait|lg|short|mio|math|no_inst)" gpurun_out/r2_ncu_$name.txt | cut -c1-150; head -8 gpurun_out/r2_ncu_${name}_source_top.txt | cut -c1-200
}
cap k_lis_plane_last 'k_lis_plane<sperr_b200::Tree3DPow2>' 14 $B
cap k_lipref_emit 'k_lipref_emit' 1 $B
du -sh gpurun_out
