"""The C-ABI library loads on a CPU-only box and exports exactly what include/stz.h declares; the
weight-blob layout agrees between spec.py and csrc/stz_layout.h; the product path fails loudly
without a GPU (no CPU fallback); the library's host-side schedule / coefficient tables (stz_debug_plan, no device
needed) reproduce the oracle's sampler loop.  No device compute call is made here."""
import ctypes as C
import os
import re

import pytest
import torch

import styletts_zs_b200 as stz
from styletts_zs_b200 import path as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return stz.load_library()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "stz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _header_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"libstz.so lacks {name}"
    assert sorted(P.EXPORTED_SYMBOLS) == declared, "path.py binds a different set than include/stz.h declares"
    assert lib.stz_abi_version() == stz.ABI_VERSION


def test_library_exports_nothing_the_header_does_not_declare():
    """The other direction: every stz_* symbol in the dynamic symbol table is declared in include/stz.h."""
    import shutil
    import subprocess
    if shutil.which("nm") is None:
        pytest.skip("nm not available")
    so = os.path.join(ROOT, "styletts-zs_b200", "csrc", "libstz.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], check=True, capture_output=True, text=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("stz_")})
    assert exported == _header_functions()


@pytest.mark.parametrize("cfg", [stz.DEFAULT, stz.TINY], ids=["default", "tiny"])
def test_weight_layout_matches_spec(lib, cfg):
    cc = P._cconfig(cfg)
    offs = stz.weight_offsets(cfg)
    assert lib.stz_weights_nfloats(C.byref(cc)) == offs["__total__"][0]
    for name, (off, shape) in offs.items():
        if name == "__total__":
            continue
        assert lib.stz_weight_offset(C.byref(cc), name.encode()) == off, name
        assert off % 64 == 0
    assert lib.stz_weight_offset(C.byref(cc), b"no.such.entry") == -1
    blob = stz.init_weights(cfg, 0)
    assert blob.numel() == offs["__total__"][0] and bool(torch.isfinite(blob).all())
    assert torch.equal(blob, stz.init_weights(cfg, 0))            # deterministic
    assert not torch.equal(blob, stz.init_weights(cfg, 1))


def test_config_struct_matches_dataclass():
    names = [f[0] for f in P._CConfig._fields_]
    assert names == list(stz.DEFAULT.as_dict().keys())
    hdr = open(os.path.join(ROOT, "include", "stz.h")).read()
    body = re.search(r"typedef struct stz_config \{(.*?)\} stz_config;", hdr, re.S).group(1)
    hdr_names = re.findall(r"\b([a-z_]+)\s*[,;]", body)
    assert hdr_names == names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    with pytest.raises(stz.StzError):
        stz.StyleTTSZSPath(stz.TINY, stz.init_weights(stz.TINY, 0))
    cc = P._cconfig(stz.DEFAULT)
    w = stz.init_weights(stz.DEFAULT, 0)
    h = C.c_void_p()
    rc = lib.stz_create(C.byref(cc), C.c_void_p(w.data_ptr()), w.numel(), 0, C.byref(h))
    assert rc == -3 and not h.value                                # STZ_E_DEVICE
    assert b"no CPU fallback" in lib.stz_last_error(None)
    with pytest.raises(stz.StzError):
        stz.StyleTTSZSPath(stz.DEFAULT, backend="oracle")


def test_create_rejects_bad_arguments(lib):
    cc = P._cconfig(stz.DEFAULT)
    h = C.c_void_p()
    assert lib.stz_create(None, None, 0, 0, C.byref(h)) == -1       # STZ_E_ARG
    bad = P._cconfig(stz.StzConfig(n_heads=7))                       # d_model != 64 * heads
    w = torch.zeros(16)
    assert lib.stz_create(C.byref(bad), C.c_void_p(w.data_ptr()), 16, 0, C.byref(h)) == -2   # STZ_E_SHAPE
    assert lib.stz_sample_style(None, None, None, None, None, None, 1, 1, 1, 1, 1.0, 0, None, None) == -1
    assert lib.stz_predict_duration(None, None, None, None, 1, 1, None, None, None) == -1


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "styletts-zs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


# ---------------------------------------------------------------------------------------------
# host logic of the library without a GPU: the schedule / coefficient tables it uploads (stz_debug_plan)
# against the oracle's schedule algebra and sampler loop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sampler,steps", [("student", 1), ("student", 4), ("student", 7), ("teacher", 1), ("teacher", 3),
                                           ("teacher", 32)])
def test_library_sampler_plan_matches_oracle_loop(sampler, steps):
    import math
    import torch
    from oracle import schedule as S
    from oracle.model import sample_loop, SAMPLER_STUDENT, SAMPLER_TEACHER
    cfg = stz.DEFAULT
    plan = stz.sampler_plan(cfg, steps, sampler, cfg_scale=1.7)
    sig, coef = plan["sigma"], plan["coef"].double()
    kind = SAMPLER_TEACHER if sampler == "teacher" else SAMPLER_STUDENT
    # a-1: the sigmas fed to the denoiser
    if sampler == "student":
        want = S.student_sigmas(steps, cfg)[:-1]
    else:
        ks = S.teacher_sigmas(steps, cfg)
        want = []
        for i in range(steps):
            want += [ks[i], S.adpm2_sigmas(ks[i], ks[i + 1])[2]]
    assert torch.allclose(sig, torch.tensor(want, dtype=torch.float64), rtol=1e-12)
    assert abs(plan["sigma0"] - want[0]) < 1e-12 and abs(plan["cin0"] - S.edm_precond(want[0], cfg.sigma_data)[2]) < 1e-12
    assert bool((coef[:, 5] == torch.tensor(1.7, dtype=torch.float32).double()).all())
    # time features of c_noise = ln(sigma) / 4
    for e, sg in enumerate(want):
        tf = torch.tensor(S.time_features(math.log(sg) / 4.0, cfg.d_time), dtype=torch.float64)
        assert torch.allclose(plan["tfeat"][e].double(), tf, atol=1e-6)
    # a-2 / a-6: drive the oracle's sampler loop and the library's fused affine updates with the same toy network
    g = torch.Generator().manual_seed(steps)
    ns = steps + 1 if sampler == "teacher" else 1
    noise = torch.randn(ns, 3, 5, generator=g, dtype=torch.float64)
    net = lambda x_in, c_noise: torch.tanh(0.3 * x_in + 0.1 * c_noise)

    def denoise(x, sigma):
        c_skip, c_out, c_in, c_noise = S.edm_precond(sigma, cfg.sigma_data)
        return c_skip * x + c_out * net(c_in * x, c_noise)
    ref = sample_loop(cfg, denoise, noise, steps, kind)
    x = plan["sigma0"] * noise[0]
    x_mid = torch.zeros_like(x)
    x_in = plan["cin0"] * x
    for e in range(len(want)):
        F = net(x_in, math.log(want[e]) / 4.0)
        cx, cm, cF, cn, cin_next, _, dest, _ = (float(v) for v in coef[e])
        if dest == 1.0:                                   # ADPM2 half step: only x_mid moves
            x_mid = cx * x + cF * F
            x_in = cin_next * x_mid
        else:
            nz = noise[e // 2 + 1] if sampler == "teacher" else 0.0
            x = cx * x + cm * x_mid + cF * F + cn * nz
            x_in = cin_next * x
        nxt = want[e + 1] if e + 1 < len(want) else None
        if nxt is not None:                               # the epilogue also emits c_in(next sigma) * state
            assert abs(cin_next - S.edm_precond(nxt, cfg.sigma_data)[2]) < 1e-6 * max(1.0, cin_next)
    assert torch.allclose(x, ref, rtol=2e-6, atol=2e-6)   # coefficients are stored in fp32


def test_header_is_plain_c_and_a_c_client_links_and_runs(tmp_path):
    """include/stz.h compiles as C99 and a C program (tests/c/abi_smoke.c: no CUDA, no C++) links against libstz.so and
    calls the host-only entry points — the boundary is a C ABI, not a C++ or torch one."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    inc, libdir = os.path.join(ROOT, "include"), os.path.join(ROOT, "styletts-zs_b200", "csrc")
    stz.load_library()                                   # builds libstz.so if it is missing
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "stz.h")], check=True)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc,
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lstz",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok 8 evaluations")
