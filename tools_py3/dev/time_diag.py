"""Device time of one output step (pic1dp_gpu_output_all) at the bench state: 1e8 markers, 64 x 64 histogram."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pic1dp_b200 as P
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = P.Pic1dGpu(P.default_params(nx=1024, capacity=n))
g.load_markers_counter(0, n, 7, 0, n)
g.collect_charge(); g.solve_field(); g.step(1)
g.output_all(64, 64, 8.0)
ts = []
for _ in range(5):
    g.timer_start(); g.output_all(64, 64, 8.0); ts.append(g.timer_stop())
print("output_all ms: min %.3f median %.3f" % (min(ts), sorted(ts)[2]), "lib", os.environ.get("PIC1DP_B200_LIB", "tree"))
