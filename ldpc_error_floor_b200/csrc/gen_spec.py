#!/usr/bin/env python
"""Build-time generator of the graph-specialised packed kernels.

For every base graph in ../data/base_graphs.npz (the reference's shipped BaseGraph/*.txt, stored as
arrays) and every launch geometry listed in GEOMETRIES (or picked by the same heuristic the host
uses), emit gen/spec_<name>.cu -- a struct of constexpr tables + one __global__ instantiation of
nms_decode_body<H2SpecPolicy<G>> -- and gen/spec_registry.cu, the table the host looks up by graph
hash at decoder creation.  Unknown graphs use the degree-bucketed generic kernels.

    python gen_spec.py [outdir]      # prints the generated file names, one per line
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "data", "base_graphs.npz")
MISC_WORDS = 400
WSTAGE_MAX = 4096
MAX_SMEM = 227 * 1024

# (graph key) -> list of (Fp, R); empty / missing = heuristic choice.  Tuned on B200 (profiles/).
GEOMETRIES = {
    # (Fp, R) per graph, measured on B200 with tools/geom_sweep.py (profiles/r01_geometry_sweep.md): few fat warps
    # (R = 2: each warp owns half of the rows / columns, 80-110 registers) beat many thin ones.  More than one
    # entry = variants compiled side by side, first = default, LDPC_B200_FP / LDPC_B200_R select at run time.
    "wimax": [(4, 2), (8, 2)],   # round-2 re-sweep (profiles/r02_geometry_sweep_wimax.txt): (8,2) 46.2, (4,2) 43.6, (16,2) 43.7, (12,2) 41.3, (8,3) 40.6, (4,3) 39.3
    "wifi": [(7, 2)],
    "5g_r073_z72": [(3, 2)],
    "5g_r050_z64": [(2, 2)],
    "5g_r050_z32": [(4, 2)],
    "5g_r033_z32": [(4, 2)],
    "5g_r073_z32": [(4, 2)],
    "mackay": [(32, 8)],
    "bch": [(32, 4)],
}
# packed kernels: the variant used for launches WITHOUT early termination when it differs from the default.  With every
# frame running all iterations fewer, fatter CTAs win (WiMAX 8,2: 16 frames per CTA, +5 %); with early termination the
# slowest frame of a CTA holds the others back, so the default stays at 8 frames per CTA (profiles/r01_geometry_sweep.md).
GEOMETRIES_NOET = {"wimax": (8, 2)}
# float32 kernels (one frame per lane): same lanes as the packed choice
GEOMETRIES_F32 = {
    "wimax": [(4, 2)],   # round-2 re-sweep (profiles/r02_f32_geometry_sweep.txt): (8,2) 19.1, (4,2) 18.8, (16,2) 18.8, (8,3) 17.1, (4,3) 16.8, (2,2) 12.6 "wifi": [(7, 2)], "5g_r073_z72": [(4, 2)], "5g_r050_z64": [(2, 2)], "5g_r050_z32": [(4, 2)],
    "5g_r033_z32": [(4, 2)], "5g_r073_z32": [(4, 2)], "mackay": [(32, 8)], "bch": [(32, 4)],
}
# persistent-slot Monte-Carlo kernels (nms_mcp.cuh: no float channel array, no ballots -> smaller CTAs, more of them per SM).
# First entry = default; LDPC_B200_MCP_FP / LDPC_B200_MCP_R select another compiled one at run time.  z72 on B200 at 5.5 dB
# (profiles/r02_mc_sweep.txt): (4,2) 28.4, (2,2) 27.7, (3,2) 26.9, (4,3) 25.9 M frames/s -- 288 = 9 x 32 lanes: no padding
# lanes, no bank conflicts at the rotation wrap.
GEOMETRIES_MCP = {
    "wimax": [(4, 2)], "wifi": [(7, 2)], "5g_r073_z72": [(4, 2), (2, 2)], "5g_r050_z64": [(2, 2)],
    "5g_r050_z32": [(4, 2)], "5g_r033_z32": [(4, 2)], "5g_r073_z32": [(4, 2)], "mackay": [(32, 8)], "bch": [(32, 4)],
}


def mcp_misc_words(FB):
    return 224 + FB * 16


def mcp_np():
    """NMS_MCP_NP of nms_common.cuh: producer warps of the persistent-slot kernels"""
    import re
    m = re.search(r"#define NMS_MCP_NP (\d+)", open(os.path.join(HERE, "nms_common.cuh")).read())
    return int(m.group(1))
SKIP = {"polar"}   # row degree 64 > 32: generic two-pass kernel


def fnv1a(M, N, z, proto):
    h = 0xcbf29ce484222325
    for v in [M, N, z] + [int(x) for x in proto.reshape(-1)]:
        for b in int(v & 0xffffffff).to_bytes(4, "little"):
            h ^= b
            h = (h * 0x100000001b3) & 0xffffffffffffffff
    return h


def by_degree(deg):
    """rows / columns sorted by degree (descending, stable) -- same order as degree_classes() in ldpc_capi.cu"""
    return sorted(range(len(deg)), key=lambda a: -deg[a])


def layout(E, N, LP, C, wwords):
    off = E * LP
    off = (off + 3) & ~3
    off_xa = off; off += N * LP * 2
    off_xq = off; off += N * LP
    off_hb = off; off += 2 * 2 * N * C
    off_w = off; off += wwords
    off_misc = off; off += MISC_WORDS
    return off_xa, off_xq, off_hb, off_w, off_misc, off


def heuristic(M, N, E, z):
    best, pick = -1.0, None
    for Fp in range(1, 33):
        L = z * Fp
        LP = (L + 31) & ~31
        C = LP // 32
        if C > 16:
            break
        smem = (layout(E, N, LP, C, 256)[-1] - N * LP) * 4   # without the xq array (decoders with VN weights drop it)
        if smem > MAX_SMEM:
            break
        R = 1
        while R * C <= 24 and R <= max(M, N):
            W = C * R
            if W >= 2:
                cps = min(MAX_SMEM // smem, 32 // W)   # 64 registers -> 32 warps per SM
                if cps >= 1:
                    lane = L / LP
                    bal = 0.45 * M / (-(-M // R) * R) + 0.55 * N / (-(-N // R) * R)
                    occ = min(1.0, cps * W / 28.0)
                    score = lane * bal * occ + 1e-4 * cps * W - 1e-5 * smem / 1024
                    if score > best:
                        best, pick = score, (Fp, R)
            R += 1
    return pick


def arr(name, vals, ty="short"):
    return f"    static constexpr {ty} {name}[{max(len(vals), 1)}] = {{{', '.join(str(int(v)) for v in vals) or '0'}}};"


def emit(key, proto, z, Fp, R, outdir, f32=False, mcp=False):
    M, N = proto.shape
    row, col, shift, row_ptr = [], [], [], [0]
    for i in range(M):
        for j in range(N):
            if proto[i, j] != -1:
                row.append(i); col.append(j); shift.append(int(proto[i, j]) % z)
        row_ptr.append(len(row))
    E = len(row)
    col_ptr = [0] * (N + 1)
    for e in range(E):
        col_ptr[col[e] + 1] += 1
    for j in range(N):
        col_ptr[j + 1] += col_ptr[j]
    fill = [0] * N
    col_edge = [0] * E
    for e in range(E):
        col_edge[col_ptr[col[e]] + fill[col[e]]] = e
        fill[col[e]] += 1
    L = z * Fp
    LP = (L + 31) & ~31
    C = LP // 32
    dc = [row_ptr[i + 1] - row_ptr[i] for i in range(M)]
    dv = [col_ptr[j + 1] - col_ptr[j] for j in range(N)]
    if max(dc) > 32 or max(dv) > 16:
        return None
    cn_order, vn_order = by_degree(dc), by_degree(dv)
    vn_e = [col_edge[k] for k in range(E)]
    vn_rot = [(L - shift[col_edge[k]] * Fp) % L for k in range(E)]
    threads = C * R * 32
    smem = (layout(E, N, LP, C, 256)[-1] - N * LP) * 4   # without the xq array (decoders with VN weights drop it)
    name = f"{key}_fp{Fp}_r{R}"
    # CN phase: degree classes in descending order and how many rows of each class a slot owns (slot s: positions s, s+R, ...)
    degs_desc = sorted(set(dc), reverse=True)
    cls_cnt = [sum(1 for p in range(s_, M, R) if dc[cn_order[p]] == dg) for s_ in range(R) for dg in degs_desc]
    cn_tables = arr('cn_degs_desc', degs_desc) + "\n" + arr('cn_cls_cnt', cls_cnt)
    if mcp:
        if 2 * Fp > 64:
            return None
        NP = mcp_np()
        threads += 32 * NP                                                  # decoding warps + producer warps
        ring = (2 * Fp * (((N * z + 3) & ~3) // 2) + 3) & ~3 if NP else 0    # NMS_MCP_RING_WORDS
        words = ((E * LP + 3) & ~3) + N * LP + ring + 256 + mcp_misc_words(2 * Fp)
        if words * 4 > MAX_SMEM:
            return None
        minb = max(1, min(MAX_SMEM // (words * 4 + 1024), 2048 // threads, 65536 // (threads * 56)))
        src = f"""// GENERATED by gen_spec.py -- do not edit.  Graph "{key}": {M}x{N}, z={z}, E={E}; persistent-slot Monte-Carlo geometry Fp={Fp} R={R}.
#include "../nms_mcp.cuh"

namespace nms {{
struct GM_{name} {{
    static constexpr int M = {M}, N = {N}, E = {E}, z = {z}, Fp = {Fp}, L = {L}, LP = {LP}, C = {C}, R = {R};
{arr('row_ptr', row_ptr)}
{arr('col_ptr', col_ptr)}
{arr('cn_order', cn_order)}
{arr('vn_order', vn_order)}
{arr('vn_e', vn_e)}
{arr('vn_rot', vn_rot)}
    static constexpr int NDEG = {len(sorted(set(dc)))};
{arr('cn_degs', sorted(set(dc)))}
{cn_tables}
}};

__global__ void __launch_bounds__({threads}, {minb}) nms_mcp_spec_{name}(const __grid_constant__ KParams P) {{
    McpKernel<GM_{name}>::run(P);
}}
}}   // namespace nms

extern "C" const void *nms_spec_mcp_func_{name}() {{ return (const void *)nms::nms_mcp_spec_{name}; }}
"""
        path = os.path.join(outdir, f"spec_mcp_{name}.cu")
        old = open(path).read() if os.path.exists(path) else None
        if old != src:
            open(path, "w").write(src)
        return dict(name=name, hash=fnv1a(M, N, z, proto), M=M, N=N, z=z, E=E, Fp=Fp, R=R, path=path, noet=0)
    if f32:
        # float path: msg + xa + ballots + weights + misc; the quantised twin adds the xq array
        words = E * LP + N * LP + 2 * N * C + E + 1 + E * C * (2 if L != LP else 1) + 256 + MISC_WORDS
        mb = [max(1, min(MAX_SMEM // (w * 4 + 1024), 2048 // threads, 65536 // (threads * 56))) for w in (words, words + N * LP)]
        src = f"""// GENERATED by gen_spec.py -- do not edit.  Graph "{key}": {M}x{N}, z={z}, E={E}; float32 geometry Fp={Fp} R={R}.
#include "../nms_f32_spec.cuh"

namespace nms {{
struct GF_{name} {{
    static constexpr int M = {M}, N = {N}, E = {E}, z = {z}, Fp = {Fp}, L = {L}, LP = {LP}, C = {C}, R = {R};
{arr('row_ptr', row_ptr)}
{arr('col_ptr', col_ptr)}
{arr('cn_order', cn_order)}
{arr('vn_order', vn_order)}
{arr('vn_e', vn_e)}
{arr('vn_rot', vn_rot)}
    static constexpr int NDEG = {len(sorted(set(dc)))}, DVMAX = {max(dv)};
{arr('cn_degs', sorted(set(dc)))}
{cn_tables}
}};

__global__ void __launch_bounds__({threads}, {mb[0]}) nms_f32_spec_{name}(const __grid_constant__ KParams P) {{
    nms_decode_body<F32SpecPolicy<GF_{name}, 0>>(P);
}}
__global__ void __launch_bounds__({threads}, {mb[1]}) nms_f32q_spec_{name}(const __grid_constant__ KParams P) {{
    nms_decode_body<F32SpecPolicy<GF_{name}, 1>>(P);
}}
}}   // namespace nms

extern "C" const void *nms_spec_f32_func_{name}() {{ return (const void *)nms::nms_f32_spec_{name}; }}
extern "C" const void *nms_spec_f32q_func_{name}() {{ return (const void *)nms::nms_f32q_spec_{name}; }}
"""
        path = os.path.join(outdir, f"spec_f32_{name}.cu")
        old = open(path).read() if os.path.exists(path) else None
        if old != src:
            open(path, "w").write(src)
        return dict(name=name, hash=fnv1a(M, N, z, proto), M=M, N=N, z=z, E=E, Fp=Fp, R=R, path=path, noet=0)
    # resident CTAs the kernel is compiled for: shared memory, threads, and >= 56 registers per thread
    minb = max(1, min(MAX_SMEM // (smem + 1024), 2048 // threads, 65536 // (threads * 56)))
    h = fnv1a(M, N, z, proto)
    src = f"""// GENERATED by gen_spec.py -- do not edit.  Graph "{key}": {M}x{N}, z={z}, E={E}; geometry Fp={Fp} R={R}.
#include "../nms_h2_spec.cuh"

namespace nms {{
struct G_{name} {{
    static constexpr int M = {M}, N = {N}, E = {E}, z = {z}, Fp = {Fp}, L = {L}, LP = {LP}, C = {C}, R = {R};
{arr('row_ptr', row_ptr)}
{arr('col_ptr', col_ptr)}
{arr('cn_order', cn_order)}
{arr('vn_order', vn_order)}
{arr('vn_e', vn_e)}
{arr('vn_rot', vn_rot)}
    static constexpr int NDEG = {len(sorted(set(dc)))};
{arr('cn_degs', sorted(set(dc)))}
{cn_tables}
}};

__global__ void __launch_bounds__({threads}, {minb}) nms_h2_spec_{name}(const __grid_constant__ KParams P) {{
    nms_decode_body<H2SpecPolicy<G_{name}>>(P);
}}
}}   // namespace nms

extern "C" const void *nms_spec_func_{name}() {{ return (const void *)nms::nms_h2_spec_{name}; }}
"""
    path = os.path.join(outdir, f"spec_{name}.cu")
    old = open(path).read() if os.path.exists(path) else None
    if old != src:
        open(path, "w").write(src)
    return dict(name=name, hash=h, M=M, N=N, z=z, E=E, Fp=Fp, R=R, path=path, noet=int(GEOMETRIES_NOET.get(key) == (Fp, R)))


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "gen")
    os.makedirs(outdir, exist_ok=True)
    d = dict(np.load(DATA))
    keys = sorted(k.split("/")[1] for k in d if k.endswith("/proto"))
    entries = []
    for key in keys:
        if key in SKIP:
            continue
        proto = d[f"graph/{key}/proto"].astype(np.int64)
        z = int(d[f"graph/{key}/meta"][0])
        M, N = proto.shape
        E = int((proto != -1).sum())
        geos = GEOMETRIES.get(key) or [heuristic(M, N, E, z)]
        for Fp, R in geos:
            e = emit(key, proto, z, Fp, R, outdir)
            if e:
                entries.append(e)
    entries32 = []
    for key in keys:
        if key in SKIP or key not in GEOMETRIES_F32:
            continue
        proto = d[f"graph/{key}/proto"].astype(np.int64)
        z = int(d[f"graph/{key}/meta"][0])
        for Fp, R in GEOMETRIES_F32[key]:
            e = emit(key, proto, z, Fp, R, outdir, f32=True)
            if e:
                entries32.append(e)
    entries_mcp = []
    for key in keys:
        if key in SKIP or key not in GEOMETRIES_MCP:
            continue
        proto = d[f"graph/{key}/proto"].astype(np.int64)
        z = int(d[f"graph/{key}/meta"][0])
        for Fp, R in GEOMETRIES_MCP[key]:
            e = emit(key, proto, z, Fp, R, outdir, mcp=True)
            if e:
                entries_mcp.append(e)
    reg = ["// GENERATED by gen_spec.py -- do not edit.", '#include "../nms_common.cuh"', ""]
    for tag, ents in (("f32", entries32), ("f32q", entries32), ("mcp", entries_mcp)):
        reg += [f'extern "C" const void *nms_spec_{tag}_func_{e["name"]}();' for e in ents]
        reg += ["", f"static const NmsSpecEntry g_spec_{tag}[] = {{"]
        reg += [f'    {{"{e["name"]}", 0x{e["hash"]:016x}ull, {e["M"]}, {e["N"]}, {e["z"]}, {e["E"]}, {e["Fp"]}, {e["R"]}, '
                f'nms_spec_{tag}_func_{e["name"]}, 0}},' for e in ents]
        reg += ["    {nullptr, 0ull, 0, 0, 0, 0, 0, 0, nullptr, 0}", "};", "",
                f'extern "C" const NmsSpecEntry *nms_spec_{tag}_table(int *count) {{',
                f"    if (count) *count = {len(ents)};", f"    return g_spec_{tag};", "}", ""]
    reg += [f'extern "C" const void *nms_spec_func_{e["name"]}();' for e in entries]
    reg += ["", "static const NmsSpecEntry g_spec[] = {"]
    reg += [f'    {{"{e["name"]}", 0x{e["hash"]:016x}ull, {e["M"]}, {e["N"]}, {e["z"]}, {e["E"]}, {e["Fp"]}, {e["R"]}, '
            f'nms_spec_func_{e["name"]}, {e["noet"]}}},' for e in entries]
    reg += ["    {nullptr, 0ull, 0, 0, 0, 0, 0, 0, nullptr, 0}", "};", "",
            'extern "C" const NmsSpecEntry *nms_spec_table(int *count) {',
            f"    if (count) *count = {len(entries)};", "    return g_spec;", "}", ""]
    rpath = os.path.join(outdir, "spec_registry.cu")
    txt = "\n".join(reg)
    if not os.path.exists(rpath) or open(rpath).read() != txt:
        open(rpath, "w").write(txt)
    for e in entries + entries32 + entries_mcp:
        print(os.path.basename(e["path"]))
    print("spec_registry.cu")


if __name__ == "__main__":
    main()
