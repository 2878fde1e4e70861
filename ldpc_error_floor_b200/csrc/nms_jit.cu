// nms_jit.cu -- run-time graph specialisation.
//
// The reference compiles ANY proto matrix at run time (Main_Functions.init_connecting_matrix, :46-150).  The build ships
// unrolled kernels for the base graphs it knows (gen_spec.py); for every other graph this file emits the same source at
// decoder creation -- constexpr tables of the graph + one instantiation of the packed decode kernel (nms_h2_spec.cuh) or
// of the persistent-slot Monte-Carlo kernel (nms_mcp.cuh) --, compiles it with NVRTC for sm_100a, keeps the cubin in an
// on-disk cache keyed by a hash of the generated source and the kernel headers, and loads it through the driver API.
// libnvrtc and the driver entry points are resolved lazily (dlopen / cudaGetDriverEntryPoint): the library still loads,
// and serves every shipped graph, where they are missing.
#include "nms_jit.h"
#include "nms_common.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

namespace {

// ---- NVRTC, resolved at first use
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *h = nullptr;
    int (*create)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*compile)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*log_size)(nvrtcProgram, size_t *) = nullptr;
    int (*log)(nvrtcProgram, char *) = nullptr;
    int (*cubin_size)(nvrtcProgram, size_t *) = nullptr;
    int (*cubin)(nvrtcProgram, char *) = nullptr;
    int (*destroy)(nvrtcProgram *) = nullptr;
    bool ok = false;
};

Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            n.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) return;
#define SYM(field, sym) n.field = (decltype(n.field))dlsym(n.h, sym)
        SYM(create, "nvrtcCreateProgram"); SYM(compile, "nvrtcCompileProgram"); SYM(log_size, "nvrtcGetProgramLogSize");
        SYM(log, "nvrtcGetProgramLog"); SYM(cubin_size, "nvrtcGetCUBINSize"); SYM(cubin, "nvrtcGetCUBIN");
        SYM(destroy, "nvrtcDestroyProgram");
#undef SYM
        n.ok = n.create && n.compile && n.log_size && n.log && n.cubin_size && n.cubin && n.destroy;
    });
    return n;
}

// ---- driver API entry points through the runtime (no link-time dependency on libcuda)
struct Driver {
    CUresult (*moduleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*occupancy)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                             void **) = nullptr;
    bool ok = false;
};

Driver &driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        d.ok = get("cuModuleLoadData", (void **)&d.moduleLoadData) && get("cuModuleGetFunction", (void **)&d.moduleGetFunction) &&
               get("cuFuncSetAttribute", (void **)&d.funcSetAttribute) &&
               get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void **)&d.occupancy) &&
               get("cuLaunchKernel", (void **)&d.launchKernel);
        if (!d.ok) cudaGetLastError();
    });
    return d;
}

unsigned long long fnv1a(const std::string &s, unsigned long long h = 0xcbf29ce484222325ull) {
    for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
    return h;
}

std::string read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string lib_dir() {
    Dl_info info;
    if (dladdr((const void *)&nms_jit_available, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

template <class T>
std::string arr(const char *name, const std::vector<T> &v) {
    std::ostringstream o;
    o << "    static constexpr short " << name << "[" << std::max<size_t>(v.size(), 1) << "] = {";
    if (v.empty()) o << "0";
    for (size_t i = 0; i < v.size(); ++i) o << (i ? ", " : "") << (int)v[i];
    o << "};\n";
    return o.str();
}

}   // namespace

extern "C" int nms_jit_available(void) { return nvrtc().ok ? 1 : 0; }

// (Fp, R) of a graph without a measured geometry, following what the sweeps on the shipped graphs found (profiles/
// r01_geometry_sweep.md): lifted graphs run best with R = 2 (two fat warps per 32-lane chunk) and the Fp that fills the
// lanes (z * Fp close to a multiple of 32) while enough warps stay resident; z = 1 graphs interleave as many frames as
// the kernel family allows and split the rows over about E / (1.5 max_dc) slots (the longest row bounds a phase).
void nms_jit_pick_geometry(int M, int N, int E, int z, int kind, int max_dc, int *Fp_out, int *R_out) {
    const int max_smem = 227 * 1024, fp_max = 32;
    const int regs = std::min(128, std::max(56, (2 * max_dc + 40 + 7) & ~7));
    auto smem_of = [&](int LP, int C) {
        const int words = kind == NMS_JIT_MCP ? ((E * LP + 3) & ~3) + N * LP + 256 + NMS_MCP_MISC_WORDS(64) + NMS_MCP_RING_WORDS(LP / z * 2, N * z) + 4
                                              : ((E * LP + 3) & ~3) + N * LP * 2 + 4 * N * C + 256 + NMS_MISC_WORDS;
        return words * 4;
    };
    if (z == 1) {
        int Fp = fp_max;
        while (Fp > 1 && smem_of((Fp + 31) & ~31, ((Fp + 31) & ~31) / 32) > max_smem) Fp /= 2;
        *Fp_out = Fp;
        *R_out = std::max(2, std::min({8, std::max(M, N), (int)(E / (1.5 * std::max(max_dc, 1)) + 0.5)}));
        return;
    }
    double best = -1.0;
    *Fp_out = 1; *R_out = 2;
    for (int Fp = 1; Fp <= fp_max; ++Fp) {
        const int L = z * Fp, LP = (L + 31) & ~31, C = LP / 32, W = 2 * C;
        if (C > 16) break;
        const int smem = smem_of(LP, C);
        if (smem > max_smem) break;
        const int cps = std::min({max_smem / (smem + 1024), 2048 / (W * 32), 65536 / (W * 32 * regs)});
        if (cps < 1) continue;
        const double score = (double)L / LP * std::min(1.0, cps * W / 20.0) - 1e-4 * smem / 1024.0;
        if (score > best) { best = score; *Fp_out = Fp; }
    }
}

// the translation unit gen_spec.py would have written for this graph
std::string nms_jit_source(const int *proto, int M, int N, int z, int Fp, int R, int kind) {
    std::vector<int> row, col, shift, row_ptr{0};
    for (int i = 0; i < M; ++i) {
        for (int j = 0; j < N; ++j) {
            const int p = proto[(size_t)i * N + j];
            if (p == -1) continue;
            row.push_back(i); col.push_back(j); shift.push_back(((p % z) + z) % z);
        }
        row_ptr.push_back((int)row.size());
    }
    const int E = (int)row.size();
    std::vector<int> col_ptr(N + 1, 0), col_edge(E), fill(N, 0);
    for (int e = 0; e < E; ++e) col_ptr[col[e] + 1]++;
    for (int j = 0; j < N; ++j) col_ptr[j + 1] += col_ptr[j];
    for (int e = 0; e < E; ++e) col_edge[col_ptr[col[e]] + fill[col[e]]++] = e;
    const int L = z * Fp, LP = (L + 31) & ~31, C = LP / 32, threads = C * R * 32 + (kind == NMS_JIT_MCP ? NMS_MCP_NP * 32 : 0);
    std::vector<int> dc(M), dv(N), cn_order(M), vn_order(N);
    for (int i = 0; i < M; ++i) dc[i] = row_ptr[i + 1] - row_ptr[i];
    for (int j = 0; j < N; ++j) dv[j] = col_ptr[j + 1] - col_ptr[j];
    std::iota(cn_order.begin(), cn_order.end(), 0);
    std::iota(vn_order.begin(), vn_order.end(), 0);
    std::stable_sort(cn_order.begin(), cn_order.end(), [&](int a, int b) { return dc[a] > dc[b]; });
    std::stable_sort(vn_order.begin(), vn_order.end(), [&](int a, int b) { return dv[a] > dv[b]; });
    std::vector<int> vn_e(E), vn_rot(E);
    for (int k = 0; k < E; ++k) { vn_e[k] = col_edge[k]; vn_rot[k] = (L - shift[col_edge[k]] * Fp) % L; }
    std::vector<int> degs(dc.begin(), dc.end());
    std::sort(degs.begin(), degs.end());
    degs.erase(std::unique(degs.begin(), degs.end()), degs.end());
    std::vector<int> degs_desc(degs.rbegin(), degs.rend()), cls_cnt;
    for (int s = 0; s < R; ++s)
        for (int dg : degs_desc) {
            int n = 0;
            for (int p = s; p < M; p += R) n += dc[cn_order[p]] == dg;
            cls_cnt.push_back(n);
        }
    const int max_smem = 227 * 1024;
    const int words = kind == NMS_JIT_MCP ? ((E * LP + 3) & ~3) + N * LP + 256 + NMS_MCP_MISC_WORDS(2 * Fp) + NMS_MCP_RING_WORDS(2 * Fp, N * z) + 4
                                          : ((E * LP + 3) & ~3) + N * LP * 2 + 4 * N * C + 256 + NMS_MISC_WORDS;
    // registers: a check row of degree dc lives in dc registers plus the tournament's temporaries
    const int max_dc = *std::max_element(dc.begin(), dc.end());
    const int regs = std::min(128, std::max(56, (2 * max_dc + 40 + 7) & ~7));
    const int minb = std::max(1, std::min({max_smem / (words * 4 + 1024), 2048 / threads, 65536 / (threads * regs)}));
    std::ostringstream o;
    o << "// generated at run time by nms_jit.cu: " << M << "x" << N << ", z=" << z << ", E=" << E << "; Fp=" << Fp << " R=" << R
      << (kind == NMS_JIT_MCP ? " (persistent-slot Monte-Carlo)" : " (packed decode)") << "\n"
      << (kind == NMS_JIT_MCP ? "#include \"nms_mcp.cuh\"\n" : "#include \"nms_h2_spec.cuh\"\n")
      << "namespace nms {\nstruct GJ {\n"
      << "    static constexpr int M = " << M << ", N = " << N << ", E = " << E << ", z = " << z << ", Fp = " << Fp << ", L = " << L
      << ", LP = " << LP << ", C = " << C << ", R = " << R << ";\n"
      << arr("row_ptr", row_ptr) << arr("col_ptr", col_ptr) << arr("cn_order", cn_order) << arr("vn_order", vn_order)
      << arr("vn_e", vn_e) << arr("vn_rot", vn_rot) << "    static constexpr int NDEG = " << degs.size() << ";\n"
      << arr("cn_degs", degs) << arr("cn_degs_desc", degs_desc) << arr("cn_cls_cnt", cls_cnt) << "};\n}   // namespace nms\n"
      << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << minb
      << ") nms_jit_kernel(const __grid_constant__ KParams P) {\n"
      << (kind == NMS_JIT_MCP ? "    nms::McpKernel<nms::GJ>::run(P);\n" : "    nms::nms_decode_body<nms::H2SpecPolicy<nms::GJ>>(P);\n")
      << "}\n";
    return o.str();
}

namespace {
// the cubin of (graph, geometry, kind): from the on-disk cache, else compiled with NVRTC and added to it
int get_cubin(const int *proto, int M, int N, int z, int Fp, int R, int kind, std::string &cubin, std::string &path, std::string &msg) {
    const std::string dir = lib_dir(), src_dir = dir + "/csrc";
    const std::string src = nms_jit_source(proto, M, N, z, Fp, R, kind);
    unsigned long long h = fnv1a(src);          // cache key: the generated source + the headers it includes
    for (const char *f : {"nms_common.cuh", "nms_device.cuh", "nms_h2.cuh", "nms_h2_spec.cuh", "nms_mcp.cuh"}) {
        const std::string t = read_file(src_dir + "/" + f);
        if (t.empty()) { msg = "kernel headers not found under " + src_dir; return -1; }
        h = fnv1a(t, h);
    }
    const char *cenv = getenv("LDPC_B200_JIT_CACHE");
    const std::string cache = cenv && *cenv ? cenv : dir + "/jit_cache";
    mkdir(cache.c_str(), 0777);
    char name[64];
    snprintf(name, sizeof name, "/%016llx.cubin", h);
    path = cache + name;
    cubin = read_file(path);
    if (!cubin.empty()) return 0;
    Nvrtc &n = nvrtc();
    if (!n.ok) { msg = "libnvrtc not available: cannot specialise this graph at run time"; return -1; }
    nvrtcProgram prog = nullptr;
    if (n.create(&prog, src.c_str(), "nms_jit.cu", 0, nullptr, nullptr) != 0) { msg = "nvrtcCreateProgram failed"; return -1; }
    const char *cuda_home = getenv("CUDA_HOME") ? getenv("CUDA_HOME") : (getenv("CUDA_PATH") ? getenv("CUDA_PATH") : "/usr/local/cuda");
    const std::string i1 = "-I" + src_dir, i2 = std::string("-I") + cuda_home + "/include";
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", i1.c_str(), i2.c_str()};
    if (n.compile(prog, 6, opts) != 0) {
        size_t ls = 0;
        n.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.log(prog, &log[0]);
        n.destroy(&prog);
        msg = "NVRTC: " + log.substr(0, 400);
        return -1;
    }
    size_t cs = 0;
    n.cubin_size(prog, &cs);
    cubin.assign(cs, '\0');
    n.cubin(prog, &cubin[0]);
    n.destroy(&prog);
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    std::ofstream out(tmp, std::ios::binary);
    out.write(cubin.data(), (std::streamsize)cubin.size());
    out.close();
    if (out) rename(tmp.c_str(), path.c_str());     // a read-only cache directory is not an error
    return 0;
}
}   // namespace

int nms_jit_build(const int *proto, int M, int N, int z, int Fp, int R, int kind, void **cufunction, char *err, int errcap) {
    auto fail = [&](const std::string &m) { if (err && errcap > 0) snprintf(err, (size_t)errcap, "%s", m.c_str()); return -1; };
    std::string cubin, path, msg;
    if (get_cubin(proto, M, N, z, Fp, R, kind, cubin, path, msg) != 0) return fail(msg);
    if (cufunction == nullptr) return 0;            // warm the cache only (no device needed)
    Driver &D = driver();
    if (!D.ok) return fail("driver entry points unavailable");
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    if (D.moduleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) return fail("cuModuleLoadData failed for " + path);
    if (D.moduleGetFunction(&fn, mod, "nms_jit_kernel") != CUDA_SUCCESS) return fail("nms_jit_kernel not found in " + path);
    *cufunction = (void *)fn;
    return 0;
}

// ---- the three things the launcher does with a kernel, for a driver-API function
int nms_jit_set_smem(void *cufunction, int bytes) {
    return driver().funcSetAttribute((CUfunction)cufunction, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes) == CUDA_SUCCESS ? 0 : -1;
}
int nms_jit_occupancy(void *cufunction, int threads, int smem, int *ctas_per_sm) {
    return driver().occupancy(ctas_per_sm, (CUfunction)cufunction, threads, (size_t)smem) == CUDA_SUCCESS ? 0 : -1;
}
int nms_jit_launch(void *cufunction, int grid, int threads, int smem, cudaStream_t st, const KParams *P) {
    void *args[] = {(void *)P};
    return driver().launchKernel((CUfunction)cufunction, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, (CUstream)st,
                                 args, nullptr) == CUDA_SUCCESS ? 0 : -1;
}
