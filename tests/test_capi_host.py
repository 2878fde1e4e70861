"""CPU: the C-ABI library loads, exports every symbol include/ldpc_b200.h declares, and its host-only
entry points (base-graph compiler) reproduce Main_Functions.init_parameter on every shipped graph.
No compute call is made: there is no GPU here and the library has no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

import ldpc_error_floor_b200 as L
from ldpc_error_floor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRAPHS = ["wimax", "wifi", "mackay", "bch", "polar", "5g_r033_z32", "5g_r050_z32", "5g_r050_z64", "5g_r073_z32",
          "5g_r073_z72"]


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    declared = set(re.findall(r"\b(ldpc_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ldpc_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "python binding and header disagree"
    assert _lib.load().ldpc_version() >= 100


def make_graph(codes, key):
    z, ps, pe, ss, se, E = (int(v) for v in codes[f"graph/{key}/meta"])
    return L.BaseGraph(codes[f"graph/{key}/proto"].astype(np.int32), z, (ps, pe), (ss, se)), E


@pytest.mark.parametrize("key", GRAPHS)
def test_graph_compiler_matches_init_parameter(codes, key):
    g, E = make_graph(codes, key)
    assert g.E == E
    assert g.rate_ref == pytest.approx(float(codes[f"graph/{key}/rate_ref"]), rel=0, abs=0)
    assert np.array_equal(g.cn_deg, codes[f"graph/{key}/cn_deg"])
    assert np.array_equal(g.vn_deg, codes[f"graph/{key}/vn_deg"])
    sig = g.sigma(codes["snr_grid"])
    assert np.allclose(sig, codes[f"graph/{key}/sigma_ref"], rtol=1e-15, atol=0)
    # E(C) edge tables: row-major order, shift = entry mod z (Main_Functions.py:69-75)
    proto = codes[f"graph/{key}/proto"]
    rows, cols = np.nonzero(proto != -1)
    assert np.array_equal(g.edge_row, rows) and np.array_equal(g.edge_col, cols)
    assert np.array_equal(g.edge_shift, proto[rows, cols] % g.z)
    assert g.info.max_dc == g.cn_deg.max() and g.info.max_dv == g.vn_deg.max()


def test_rate_quirk_and_true_rate(codes):
    g, _ = make_graph(codes, "wimax")
    assert g.rate_ref == pytest.approx(431 / 574)         # the "+1" when nothing is punctured
    assert g.rate_true == pytest.approx(0.75) and g.k_true == 432 and g.n_true == 576
    assert g.sigma([3.0], use_ref_rate=False)[0] == pytest.approx(np.sqrt(1 / (2 * 0.75 * 10 ** 0.3)))
    g5, _ = make_graph(codes, "5g_r050_z64")
    assert g5.rate_ref == 0.5 == g5.rate_true and g5.k_true == 512 and g5.n_true == 1024


def test_init_parameter_dropin(codes):
    proto = codes["graph/wimax/proto"].astype(np.int32)
    M, N, base, cn, vn, E, rate, sigma = L.init_parameter(proto, np.array([2, 2.5, 3.0, 3.5, 4.0]), 24, 0, 0, 0, 0)
    assert (M, N, E) == (6, 24, 88) and base.sum() == 88 and rate == pytest.approx(0.7508710801393729)
    assert np.allclose(sigma, [0.64818998, 0.6119308, 0.57769993, 0.5453839, 0.5148756], atol=1e-8)


def test_error_codes_not_exceptions(codes):
    lib = _lib.load()
    h = ctypes.c_void_p()
    proto = np.zeros((2, 3), np.int32)
    assert lib.ldpc_graph_create(proto.ctypes.data, 2, 3, 0, 0, 0, 0, 0, ctypes.byref(h)) == -1    # z <= 0
    assert b"bad arguments" in lib.ldpc_last_error()
    assert lib.ldpc_graph_create(proto.ctypes.data, 2, 3, 4, 5, 2, 0, 0, ctypes.byref(h)) == -1    # pe < ps
    assert lib.ldpc_graph_create(None, 2, 3, 4, 0, 0, 0, 0, ctypes.byref(h)) == -1
    with pytest.raises(L.LdpcError):
        L.BaseGraph(proto, -3)
    assert lib.ldpc_graph_info(None, None) == -1
    assert lib.ldpc_decode(None, None, 1, 0, 0, None, 0, None, None, None, None, None) == -1


def test_no_cpu_fallback(codes):
    """Without a CUDA device the product path must fail loudly, not decode on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g, _ = make_graph(codes, "wimax")
    ws = L.WeightSet([3, 3, 3], {i: codes[f"weights/wimax_base20/block{i}"] for i in range(3)})
    with pytest.raises(L.LdpcError) as e:
        L.NMSDecoder(g, ws)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_decoder_argument_checks(codes):
    """check_params rules (Main_Functions.py:515-521) are enforced by ldpc_decoder_create before any CUDA call."""
    lib = _lib.load()
    g, _ = make_graph(codes, "wimax")
    w = np.ones((5, 88), np.float32)
    out = ctypes.c_void_p()

    def create(sharing, dt=2, qb=5, T=5, clip=20.0):
        sh = (ctypes.c_int32 * 3)(*sharing)
        return lib.ldpc_decoder_create(g._h, sh, T, w.ctypes.data, w.ctypes.data, w.ctypes.data, dt, qb,
                                       ctypes.c_float(clip), 0, ctypes.byref(out))
    assert create([3, 3, 1]) == -1 and b"sharing[2]" in lib.ldpc_last_error()
    assert create([3, 2, 3]) == -1 and b"sharing[1]" in lib.ldpc_last_error()
    assert create([4, 0, 3]) == -3                      # temporal sharing passes the checks (no GPU here -> LDPC_E_CUDA)
    assert create([5, 0, 3]) == -2                      # code 5 has no branch in build_neural_network
    assert create([3, 0, 4]) == -1 and b"sharing[2]" in lib.ldpc_last_error()
    assert create([3, 3, 3], dt=0) == -3                # sum-product passes the checks (no GPU here -> LDPC_E_CUDA)
    assert create([3, 3, 3], dt=3) == -2                # decoding_type 3: undocumented in the reference, not offered
    assert create([3, 3, 3], qb=7) == -1
    assert create([3, 3, 3], T=0) == -1
    assert create([3, 3, 3], clip=0.0) == -1


def test_host_logic_helpers():
    from ldpc_error_floor_b200 import decoder as D
    import torch
    snr = D.check_params(1, np.array([2.0, 3.0]), [3, 3, 3], 30, 20, 10)
    assert np.array_equal(snr, [0.0])                   # sampling_type 1 collapses the SNR list (:499-501)
    with pytest.raises(ValueError):
        D.check_params(2, np.array([2.0, 3.0]), [3, 0, 3], 20, 0, 20)
    with pytest.raises(ValueError):
        D.check_params(0, np.array([2.0]), [0, 0, 0], 20, 0, 20)
    with pytest.raises(ValueError):
        D.check_params(0, np.array([2.0]), [4, 0, 3], 25, 0, 20)
    with pytest.raises(ValueError):
        D.check_params(0, np.array([2.0]), [3, 0, 1], 20, 0, 20)
    with pytest.raises(ValueError):
        D.check_params(0, np.array([2.0]), [3, 2, 3], 20, 0, 20)
    packed = torch.tensor([[0b1011, 1 << 31], [0, 5]], dtype=torch.int64).to(torch.int32)
    bits = D.unpack_bits(packed, 40)
    assert bits.shape == (2, 40) and bits[0, :4].tolist() == [1, 1, 0, 1] and int(bits[1, 32]) == 1


def test_snr_point_statistics():
    from ldpc_error_floor_b200.montecarlo import SnrPoint
    pt = SnrPoint(3.0, 0.577, bits_per_frame=576)
    pt.add([1000, 55, 54, 1456, 20000, 55, 0, 54])
    assert pt.fer == 0.054 and pt.fer_last == 0.055 and pt.avg_iters == 20.0
    assert pt.ber_last == pytest.approx(1456 / 576000)
    lo, hi = pt.fer_ci95()
    assert lo < 0.054 < hi and hi - lo < 0.03


def test_two_stage_heuristic():
    """MonteCarlo's choice of the stage-1 length (montecarlo._pick_stage1) on the statistics measured on B200
    (profiles/r01_two_stage_mc.txt): on for the 5G n2112 floor region and WiMAX at 3 dB, off where stragglers are rare or
    convergence is slow."""
    from ldpc_error_floor_b200 import montecarlo

    def cnt(f, avg, frames=1e6):
        c = np.zeros(8)
        c[0], c[5], c[4] = frames, f * frames, avg * frames
        return c
    assert [montecarlo._pick_stage1(cnt(f, a), 20) for f, a in ((.128, 8.13), (.116, 7.11), (.101, 6.26), (.083, 5.53))] == [9, 8, 7, 6]
    assert montecarlo._pick_stage1(cnt(.057, 7.18), 20) == 9          # WiMAX 3 dB
    assert montecarlo._pick_stage1(cnt(.017, 8.78), 20) == 0          # 5G n1024: stragglers too rare
    assert montecarlo._pick_stage1(cnt(.0002, 3.18), 20) == 0         # WiMAX 4 dB
    assert montecarlo._pick_stage1(cnt(.7, 18.0), 20) == 0            # waterfall: most frames fail
    assert montecarlo._pick_stage1(np.zeros(8), 20) == 0


def test_jit_prebuild_needs_no_gpu(tmp_path, monkeypatch):
    """Run-time specialisation cross-compiles with NVRTC: the cubin of an unknown graph lands in the on-disk cache
    without a device (ldpc_jit_prebuild); a second call is a cache hit."""
    import time
    monkeypatch.setenv("LDPC_B200_JIT_CACHE", str(tmp_path))
    proto = -np.ones((3, 6), dtype=np.int32)
    proto[0, [0, 1, 3]] = [1, 2, 0]; proto[1, [1, 2, 3, 4]] = [3, 0, 5, 0]; proto[2, [0, 2, 4, 5]] = [4, 6, 1, 0]
    n = _lib.jit_prebuild(proto, 8)
    assert n == 2                                                  # packed decode kernel + persistent-slot Monte-Carlo kernel
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 2 and all(f.endswith(".cubin") for f in files)
    t0 = time.time()
    assert _lib.jit_prebuild(proto, 8) == 2 and time.time() - t0 < 2.0


# ---------------------------------------------------------------- host side of the int8 transport (csrc/host_pack.cpp)
def _ref_pack(x, step, qmax, lossless):
    """numpy restatement: Q(x) / step as the kernels compute it (nms_device.cuh qf: clamp to +-1e5, (x + M) - M, clamp to
    +-qmax; Main_Functions.py:475-494), or the exact 'is on the grid and in range' test."""
    qk = np.float32(1.0 / step)
    kmax = qmax / step
    x = x.astype(np.float32)
    with np.errstate(all="ignore"):
        if not lossless:
            v = np.where(x > -1e5, x, np.float32(-1e5))          # NaN -> -bound, as fmaxf(NaN, -b)
            v = np.where(v < 1e5, v, np.float32(1e5))
            return np.clip(np.rint(v * qk), -kmax, kmax).astype(np.int8), 0
        y = x * qk
        inr = np.abs(y) <= kmax
        k = np.where(inr, np.rint(np.where(inr, y, 0)), 0)
        return k.astype(np.int8), int((~(inr & (k == y))).sum())


@pytest.mark.parametrize("step, qmax", [(0.5, 7.5), (1.0, 15.0), (1.0, 7.0), (2.0, 6.0)])
def test_pack_q8_values_matches_the_quantiser(step, qmax):
    """ldpc_pack_q8_values == Q() of the reference (every q_bit that has an int8 form), quirk values included; the lossless
    verdict counts exactly the values that are off the grid or out of range; odd lengths, several pool blocks."""
    from ldpc_error_floor_b200 import _lib
    rng = np.random.default_rng(int(step * 10 + qmax))
    quirks = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-4, -1e-4, 0.25, 0.75, -0.25, 1e30, -1e30, 7.5, -7.5, 7.75,
                       8.0, 15.0, 15.5, -15.5, 6.0, 1e5, -1e5, 2e5, 0.24999999, 2.5, 3.5, -2.5, -3.5], dtype=np.float32)
    for n in (0, 1, 31, 32, 33, 1000, 65536 * 3 + 17, 1_000_003):
        x = (rng.standard_normal(n) * 6).astype(np.float32)
        if n >= quirks.size:
            x[:quirks.size] = quirks
        q, bad = _lib.pack_q8_values(x, step, qmax, False)
        r, _ = _ref_pack(x, step, qmax, False)
        assert bad == 0 and np.array_equal(q, r)
        q, bad = _lib.pack_q8_values(x, step, qmax, True)
        _, rb = _ref_pack(x, step, qmax, True)
        assert bad == rb
        xg = (np.clip(np.rint(np.nan_to_num(x, posinf=0, neginf=0) / step), -qmax / step, qmax / step) * step).astype(np.float32)
        q, bad = _lib.pack_q8_values(xg, step, qmax, True)
        assert bad == 0 and np.array_equal(q.astype(np.float32) * np.float32(step), xg)


def test_pack_q8_values_scalar_and_vector_bodies_agree(monkeypatch):
    """The AVX2 body and the portable one give the same bytes (LDPC_B200_NO_AVX2 is read once per process: run the portable
    body in a child)."""
    import subprocess
    import sys
    code = ("import numpy as np, sys; sys.path.insert(0, %r); from ldpc_error_floor_b200 import _lib; "
            "x = (np.random.default_rng(5).standard_normal(200003) * 6).astype(np.float32); x[:4] = [np.nan, -0.0, 1e30, 0.25]; "
            "q, b = _lib.pack_q8_values(x, 0.5, 7.5, False); q2, b2 = _lib.pack_q8_values(x, 0.5, 7.5, True); "
            "ok = np.abs(x * 2) <= 15; ok &= np.rint(np.where(ok, x * 2, 0)) == x * 2; "      # bytes of unencodable values are unspecified
            "import hashlib; print(hashlib.sha1(q.tobytes() + q2[ok].tobytes()).hexdigest(), b, b2)" % ROOT)
    outs = []
    for env in ({}, {"LDPC_B200_NO_AVX2": "1"}, {"LDPC_B200_HOST_THREADS": "1"}):
        e = dict(os.environ); e.update(env)
        outs.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, check=True).stdout.strip())
    assert outs[0] == outs[1] == outs[2], outs


def test_pack_q8_values_rejects_bad_arguments():
    from ldpc_error_floor_b200 import _lib
    x = np.zeros(8, dtype=np.float32)
    for step, qmax in ((0.0, 7.5), (0.5, 0.0), (0.01, 7.5)):       # qmax / step must fit int8
        with pytest.raises(_lib.LdpcError):
            _lib.pack_q8_values(x, step, qmax, False)


def test_no_fused_packed_multiply_add_in_the_decode_kernels():
    """The quantiser rounds twice (float32 product, then rint to the grid); ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into
    one FFMA2 behind the programmer's back, which rounds once and differs for weights like 0.9f (nms_device.cuh
    mul2_rn_unfused).  Static guard on the built objects: every FFMA2 in the packed decode kernels has a zero addend."""
    import shutil
    import subprocess
    build = os.path.join(ROOT, "ldpc_error_floor_b200", "csrc", "build")
    objs = [os.path.join(build, n) for n in ("spec_wimax_fp8_r2.o", "spec_wimax_fp4_r2.o", "spec_mcp_5g_r073_z72_fp4_r2.o",
                                             "spec_mcp_wimax_fp4_r2.o", "nms_h2_16_8.o", "nms_h2_0_0.o")]
    objs = [o for o in objs if os.path.exists(o)]
    if not objs or shutil.which("cuobjdump") is None:
        pytest.skip("no built objects / cuobjdump here")
    seen = 0
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, check=True).stdout
        for line in sass.splitlines():
            if "FFMA2" in line:
                seen += 1
                assert "RZ.F32" in line.split(";")[0], f"{os.path.basename(o)}: fused packed multiply-add: {line.strip()}"
    assert seen > 0


def test_host_pool_survives_a_fork():
    """A forked child (Python multiprocessing) inherits the pool object but none of its threads; it must get its own."""
    import subprocess
    import sys
    code = ("import os, sys, numpy as np; sys.path.insert(0, %r); from ldpc_error_floor_b200 import _lib; "
            "x = (np.random.default_rng(0).standard_normal(3000000) * 5).astype(np.float32); "
            "q, b = _lib.pack_q8_values(x, 0.5, 7.5, False); pid = os.fork(); "
            "ok = pid != 0 or np.array_equal(q, _lib.pack_q8_values(x, 0.5, 7.5, False)[0]); "
            "pid == 0 and os._exit(0 if ok else 3); "
            "print(os.WEXITSTATUS(os.waitpid(pid, 0)[1]))" % ROOT)
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "0", (out.stdout, out.stderr)


def test_host_stats_struct_layout_matches_the_header(tmp_path):
    """ldpc_host_stats_t as the C compiler lays it out == the ctypes mirror in _lib.HostStats (field by field)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc here")
    fields = [name for name, _ in _lib.HostStats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ldpc_b200.h"\nint main(void) {\n'
                   + "".join(f'  printf("%zu\\n", offsetof(ldpc_host_stats_t, {f}));\n' for f in fields)
                   + '  printf("%zu\\n", sizeof(ldpc_host_stats_t));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [getattr(_lib.HostStats, f).offset for f in fields] + [ctypes.sizeof(_lib.HostStats)]
    assert got == want
