"""Host model of the kernels' arithmetic shortcuts (tests/csrc/fastdiv_model.c): the FMA-based division fast paths of
pic1dp_b200/csrc/particle_kernels.cuh (div_const, div_pos with the exact-residual test div_suspect) and the select form
of the periodic wrap.  On the host fma() is exact and a / b, fmod() are the IEEE results the reference's x86-64 build
computes, so this checks the claim behind the bit-exact cell index without a GPU: every operand pair the kernels do NOT
flag for the IEEE fallback gives RN(a / b) bit for bit, and ordinary operands are never flagged."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unflagged_fast_divisions_are_correctly_rounded(tmp_path):
    exe = tmp_path / "fastdiv_model"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-std=c11", "-o", str(exe),
                    os.path.join(ROOT, "tests", "csrc", "fastdiv_model.c"), "-lm"], check=True)
    for seed in (1, 2, 3):
        out = subprocess.run([str(exe), "1000000", str(seed)], capture_output=True, text=True)
        m = re.match(r"checked (\d+) flagged (\d+) mismatches (\d+) random_flagged (\d+) of (\d+)", out.stdout)
        assert m, out.stdout + out.stderr
        checked, flagged, bad, rnd_flagged, rnd = map(int, m.groups())
        assert out.returncode == 0 and bad == 0
        assert checked > 10_000_000 and rnd > 7_000_000
        assert rnd_flagged == 0            # the hot loop does not fall back on ordinary data
        assert 0 < flagged < checked // 10000   # the adversarial near-midpoint operands are caught by the residual test


def test_model_matches_the_kernel_source():
    """The model restates three device functions; keep their defining lines in step with the kernel header."""
    k = open(os.path.join(ROOT, "pic1dp_b200", "csrc", "particle_kernels.cuh")).read()
    for needle in ("const double r1 = fma(-q1, b, a);", "(1076 << 20)", "q[k] = fma(r, y, q0);",
                   "rare = rare | !(a[k] >= 0x1p-800) | div_suspect(a[k], b, q[k]);", "e = fma(e, e, e);",
                   "(eb | ea) >= (1000u << 20)", "double xw = (x[k] >= lx) ? dsub(x[k], lx) : x[k];"):
        assert needle in k, needle


def test_fast_exp_stays_near_one_ulp(tmp_path):
    """exp_fast* with the coefficients of the kernel header, against 80-bit expl(): at most ~1.05 ulp over
    [-700, 0] (two-term Cody-Waite reduction, final + 1 rounding; glibc's own exp is < 1 ulp), which is what bounds the only non-bit-exact
    quantity of the push (w, through -d ln f0 / dv) at ~1e-16 relative per substep."""
    k = open(os.path.join(ROOT, "pic1dp_b200", "csrc", "particle_kernels.cuh")).read()
    poly = re.search(r"c_exp_poly\[10\] = \{(.*?)\};", k, re.S).group(1)
    red = re.search(r"c_exp_red\[3\] = \{(.*?)\};", k, re.S).group(1)
    red = re.sub(r"/\*.*?\*/", "", red)
    (tmp_path / "exp_coeffs.h").write_text(f"static const double c_exp_poly[10] = {{{poly}}};\n"
                                           f"static const double c_exp_red[3] = {{{red}}};\n")
    exe = tmp_path / "exp_model"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-std=c11", "-I", str(tmp_path), "-o", str(exe),
                    os.path.join(ROOT, "tests", "csrc", "exp_model.c"), "-lm"], check=True)
    out = subprocess.run([str(exe), "4000000", "5"], capture_output=True, text=True).stdout
    m = re.match(r"max_ulp ([0-9.]+) at (\S+) thermal_max_ulp ([0-9.]+) differ_from_glibc (\d+) of (\d+)", out)
    assert m, out
    assert float(m.group(1)) < 1.1 and float(m.group(3)) < 1.1, out
    assert int(m.group(4)) < int(m.group(5)) // 4      # and most results coincide with glibc's exp() anyway
