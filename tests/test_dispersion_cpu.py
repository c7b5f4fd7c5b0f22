"""The analytic numbers the physics acceptance tests compare against, computed instead of quoted: roots of the kinetic
dispersion relation the reference's own analysis tool solves (/root/reference/tools/dispersion.py:130-157), restated for
py3 in tools_py3/dispersion.py."""
import re
import os

from tools_py3 import dispersion as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bump_on_tail_root_at_the_default_wavenumber():
    """Default input (src/pic1dp_input.F90:47-72): k = 0.36, bulk n = 0.9, bump 0.1 at v0 = 5, T = T2 = 1.
    SURVEY 8c quotes omega = 1.16938 + 0.083831 i for it."""
    om = D.solve_omega(0.36, D.BUMP_ON_TAIL)
    assert abs(D.dispersion_function(om, 0.36, D.BUMP_ON_TAIL)) < 1e-12
    assert abs(om.real - 1.16938) < 5e-6 and abs(om.imag - 0.083831) < 5e-7


def test_landau_damping_root_of_a_thermal_plasma():
    """configs[2]: Maxwellian, k = 0.5 -> omega = 1.4156 - 0.1533 i (the textbook Landau damping rate)."""
    om = D.solve_omega(0.5, D.THERMAL, (1.4 - 0.1j, 1.5 - 0.2j, 1.45 - 0.15j))
    assert abs(D.dispersion_function(om, 0.5, D.THERMAL)) < 1e-12
    assert abs(om.real - 1.4156) < 1e-4 and abs(om.imag + 0.1533) < 1e-4


def test_acceptance_tests_use_these_roots():
    """The constants hard-wired into the GPU physics tests and the long-run tool are these roots."""
    g = D.solve_omega(0.36, D.BUMP_ON_TAIL).imag
    for rel in ("tests/test_gpu_physics.py", "tools_py3/long_run.py", "tests/test_oracle_hotpath.py"):
        txt = open(os.path.join(ROOT, rel)).read()
        vals = [float(v) for v in re.findall(r"0\.08383\d*", txt)]
        assert vals, rel
        assert all(abs(v - g) < 1e-6 for v in vals), (rel, vals, g)
    ld = D.solve_omega(0.5, D.THERMAL, (1.4 - 0.1j, 1.5 - 0.2j, 1.45 - 0.15j))
    txt = open(os.path.join(ROOT, "tests/test_gpu_physics.py")).read()
    assert "-0.1533" in txt and "1.4156" in txt
    assert abs(ld.imag + 0.1533) < 1e-4 and abs(ld.real - 1.4156) < 1e-4


def test_zero_of_a_cold_limit():
    """Sanity of the Z-function form: for k -> 0 the thermal root tends to the plasma frequency (omega_p = 1)."""
    om = D.solve_omega(0.05, D.THERMAL, (1.0 - 0.0j, 1.01 - 0.001j, 0.99 + 0.001j))
    assert abs(om.real - (1.0 + 1.5 * 0.05 ** 2)) < 1e-4 and abs(om.imag) < 1e-6  # Bohm-Gross: 1 + 3/2 k^2


def test_traffic_json_is_what_the_converter_makes_of_the_committed_ncu_csv():
    """profiles/r02_traffic.json (the source of bench.py's roofline.traffic) must be reproducible from the ncu CSV
    committed beside it: measured DRAM bytes per marker at the bench size, within 1 % of the algorithmic 56 / 80 B."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = json.load(open(os.path.join(root, "profiles", "r02_traffic.json")))
    out = subprocess.run([sys.executable, os.path.join(root, "tools_py3", "traffic_json.py"),
                          os.path.join(root, "profiles", "r02_traffic.csv"), str(want["markers"]), str(want["nx"]),
                          str(want["deposit_mode"])], capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    for irk, alg in (("irk1", 56), ("irk2", 80)):
        assert got[irk]["dram_bytes_per_marker"] == want[irk]["dram_bytes_per_marker"]
        assert abs(got[irk]["dram_bytes_per_marker"] / alg - 1.0) < 0.01
