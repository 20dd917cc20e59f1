# Last verification of the round on HEAD: GPU tests, smoke, a short bench, the resampler's measured error
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_final.log
tail -3 $O/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
python - <<'PY'
import numpy, torch
from sidekit_b200 import synth
from sidekit_b200.nnet.preprocessor import Resample
from tests.helpers import golden
g = golden("resample.npz")
worst = 0.0
for i in range(int(g["n_cases"])):
    fo, fn, n = (int(v) for v in g["case%d_rates" % i])
    out = Resample(fo, fn)(synth.synth_wave(2, n, seed=300 + i).cuda()).cpu().numpy()
    worst = max(worst, float(numpy.abs(out - g["case%d_y" % i]).max()))
print("resample kernel vs torchaudio golden: max abs err %.3e" % worst)
PY
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err
python tools/ab_print.py $O/bench_final.json
