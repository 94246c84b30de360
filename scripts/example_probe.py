"""The example problem (configs[1]) for launch-list captures: a few RHS evaluations and Tsit5 steps."""
import sys
sys.path.insert(0, ".")
import bench
import oetqf_b200 as oq
oq.init(0)
print(bench.example_extra(oq))
print(bench.fft_form_extra(oq))
