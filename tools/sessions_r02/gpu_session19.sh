# round 2, session 19: FFT butterflies as opaque multiply-add chains (alu pipe -> fma pipe): forward only / inverse only / both
# (variants/v8_fma1/2/3.so) against the default build at 262 144 instances, interleaved twice; then the GPU suite on the fastest variant
mkdir -p gpurun_out
R=gpurun_out/s19_variants.txt
: > $R
for rep in 1 2; do
  for lib in default variants/v8_fma3.so variants/v8_fma2.so variants/v8_fma1.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib --batch 262144 --steps 24" >> $R
    timeout 120 python bench.py --batch 262144 --steps 24 --no-e2e --no-cpu-baseline --no-parity --no-realtime 2>>gpurun_out/s19_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> $R 2>&1
  done
done
unset AECM_B200_LIB
cat $R
BEST=$(python - <<'PY'
import re, collections
v = collections.defaultdict(list)
name = None
for line in open("gpurun_out/s19_variants.txt"):
    if line.startswith("=="):
        name = line.split()[1]
    else:
        m = re.match(r"([0-9.e+]+) blocks/s", line)
        if m and name:
            v[name].append(float(m.group(1)))
mean = {k: sum(x) / len(x) for k, x in v.items() if x}
best = max((k for k in mean if k != "default"), key=lambda k: mean[k], default="")
print(best)
PY
)
echo "fastest variant: $BEST" | tee -a $R
if [ -n "$BEST" ]; then
  AECM_B200_LIB=$PWD/$BEST timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s19_tests_best.log 2>&1; echo "tests ($BEST) rc=$?" >> gpurun_out/s19_tests_best.log; tail -4 gpurun_out/s19_tests_best.log
  for cfg in "--batch 65536 --fs 8000 --steps 60" "--batch 4096 --steps 200"; do
    for lib in default $BEST; do
      if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
      echo "== $lib $cfg" >> $R
      timeout 120 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity --no-realtime 2>>gpurun_out/s19_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> $R 2>&1
    done
  done
  tail -8 $R
fi
