// Straight-line code for one compiled expression program: halo2's `Evaluator::evaluate_h` specialised per constraint system at key-creation time.
//
// expr.cu evaluates a `Program` with a warp-uniform interpreter (every instruction re-decodes three words, values round-trip through shared-memory
// slots: ~25 % of the issued instructions).  Here the same program -- same DAG, same schedule, same 7 value slots -- is printed as CUDA source with the
// slots as local variables and the operands resolved at compile time, compiled by NVRTC for sm_100a once per proving key, and launched through the
// driver API.  Only the constant table (the challenges theta, beta, gamma, y^e, delta^j beta) changes per proof; it stays a device buffer.
// libnvrtc / libcuda are loaded with dlopen: the library still links nothing but cudart, loads on machines without a driver (CPU-side ABI tests), and
// falls back to the interpreter (never to the CPU) when NVRTC is not installed or SB_NO_JIT is set.
#include <dlfcn.h>

#include <cuda.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <sstream>
#include <string>

#include "prover.h"

namespace sb {

namespace {

const char *FP_CUH_SOURCE =
#include "_obj/fp_cuh_embed.inc"
    ;

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_NEG = 3, OP_COPY = 4 };
enum { K_REG = 0, K_CONST = 1, K_INPUT = 2 };

struct Api {
    bool ok = false;
    std::string why;
    decltype(&nvrtcCreateProgram) createProgram = nullptr;
    decltype(&nvrtcCompileProgram) compileProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) getCubinSize = nullptr;
    decltype(&nvrtcGetCUBIN) getCubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) getLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) getLog = nullptr;
    decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
    decltype(&cuModuleLoadData) moduleLoadData = nullptr;
    decltype(&cuModuleGetFunction) moduleGetFunction = nullptr;
    decltype(&cuModuleUnload) moduleUnload = nullptr;
    decltype(&cuLaunchKernel) launchKernel = nullptr;
    decltype(&cuFuncSetAttribute) funcSetAttribute = nullptr;
};

Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void *rt = nullptr;
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
            rt = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (rt) break;
        }
        void *drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!rt || !drv) {
            a.why = !rt ? "libnvrtc not found" : "libcuda.so.1 not found";
            return;
        }
#define SB_SYM(lib, field, sym)                                   \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(lib, sym)); \
    if (!a.field) { a.why = std::string("missing symbol ") + sym; return; }
        SB_SYM(rt, createProgram, "nvrtcCreateProgram")
        SB_SYM(rt, compileProgram, "nvrtcCompileProgram")
        SB_SYM(rt, getCubinSize, "nvrtcGetCUBINSize")
        SB_SYM(rt, getCubin, "nvrtcGetCUBIN")
        SB_SYM(rt, getLogSize, "nvrtcGetProgramLogSize")
        SB_SYM(rt, getLog, "nvrtcGetProgramLog")
        SB_SYM(rt, destroyProgram, "nvrtcDestroyProgram")
        SB_SYM(drv, moduleLoadData, "cuModuleLoadData")
        SB_SYM(drv, moduleGetFunction, "cuModuleGetFunction")
        SB_SYM(drv, moduleUnload, "cuModuleUnload")
        SB_SYM(drv, launchKernel, "cuLaunchKernel")
        SB_SYM(drv, funcSetAttribute, "cuFuncSetAttribute")
#undef SB_SYM
        a.ok = true;
    });
    return a;
}

const int JIT_THREADS = 128;
const int JIT_MAX_COLS = 64;

// what the generated kernel takes (the layout is repeated in the generated source)
struct JitArgs {
    const void *cols[JIT_MAX_COLS];
    const void *consts;
    void *out;
    unsigned long long mask;
    long long rot_mul;  // rows per unit of rotation
    unsigned char shift[JIT_MAX_COLS];
};

std::string operand_src(const Program &p, uint32_t o) {
    const uint32_t kind = o >> 30, idx = o & 0x3fffffffu;
    std::ostringstream s;
    if (kind == K_REG) s << "s" << idx;
    else if (kind == K_CONST) s << "ldg_fp<FrParams>(C + " << 2 * idx << ")";
    else s << "ld(A, " << p.inputs[2 * idx] << ", " << p.inputs[2 * idx + 1] << ", row)";
    return s.str();
}

std::string generate_source(const Program &p) {
    std::ostringstream s;
    s << "#ifdef __CUDACC_RTC__\ntypedef unsigned int uint32_t;\ntypedef unsigned long long uint64_t;\ntypedef int int32_t;\ntypedef long long int64_t;\ntypedef unsigned char uint8_t;\n"
         "#else\n#include <stdint.h>\n#endif\n";
    s << FP_CUH_SOURCE << "\n";
    s << "using namespace sb;\ntypedef Fp<FrParams> F;\n";
    s << "__device__ __noinline__ F jmul(F a, F b) { return mul(a, b); }\n";
    s << "struct JitArgs { const uint4 *cols[" << JIT_MAX_COLS << "]; const uint4 *consts; uint4 *out; unsigned long long mask; long long rot_mul; unsigned char shift["
      << JIT_MAX_COLS << "]; };\n";
    s << "__device__ __forceinline__ F ld(const JitArgs &A, int col, int rot, unsigned long long row) {\n"
         "    const unsigned long long j = (row + (unsigned long long)((long long)rot * A.rot_mul)) & A.mask;\n"
         "    return ldg_fp<FrParams>(A.cols[col] + 2 * (j << A.shift[col]));\n}\n";
    s << "extern \"C\" __global__ void __launch_bounds__(" << JIT_THREADS << ") sb_h_jit(const JitArgs A) {\n";
    s << "    const unsigned long long row = blockIdx.x * " << JIT_THREADS << "ull + threadIdx.x;\n    const uint4 *C = A.consts;\n";
    for (uint32_t i = 0; i < p.n_slots; i++) s << "    F s" << i << " = F::zero();\n";
    for (size_t pc = 0; pc < p.code.size() / 3; pc++) {
        const uint32_t w0 = p.code[3 * pc], wa = p.code[3 * pc + 1], wb = p.code[3 * pc + 2];
        const uint32_t op = w0 >> 16, dst = w0 & 0xffffu;
        s << "    s" << dst << " = ";
        if (op == OP_MUL) s << "jmul(" << operand_src(p, wa) << ", " << operand_src(p, wb) << ")";
        else if (op == OP_ADD) s << "add(" << operand_src(p, wa) << ", " << operand_src(p, wb) << ")";
        else if (op == OP_SUB) s << "sub(" << operand_src(p, wa) << ", " << operand_src(p, wb) << ")";
        else if (op == OP_NEG) s << "neg(" << operand_src(p, wa) << ")";
        else s << operand_src(p, wa);
        s << ";\n";
    }
    s << "    store_fp(A.out + 2 * row, s" << p.out_slot << ");\n}\n";
    return s.str();
}

}  // namespace

std::string expr_jit_source(const Program &prog) { return generate_source(prog); }

struct ExprJit {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    uint32_t n_consts = 0, max_col = 0;
};

void expr_jit_free(ExprJit *j) {
    if (!j) return;
    if (j->mod && api().ok) api().moduleUnload(j->mod);
    delete j;
}

// Compile `prog` (its STRUCTURE: the constant table is an argument of every launch).  Returns nullptr, with the reason in *why, when NVRTC is not usable.
ExprJit *expr_jit_compile(const Program &prog, std::string *why) {
    Api &a = api();
    if (!a.ok) { if (why) *why = a.why; return nullptr; }
    uint32_t max_col = 0;
    for (size_t i = 0; i < prog.inputs.size(); i += 2) max_col = std::max<uint32_t>(max_col, (uint32_t)prog.inputs[i]);
    if (max_col >= (uint32_t)JIT_MAX_COLS || prog.n_slots > 24) { if (why) *why = "program outside the JIT's limits"; return nullptr; }
    const std::string src = generate_source(prog);
    // one compilation per distinct program per process: every key of the same constraint system shares the cubin (NVRTC takes seconds)
    static std::mutex cache_mu;
    static std::map<std::string, std::string> cubin_cache;
    std::string cubin;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        auto it = cubin_cache.find(src);
        if (it != cubin_cache.end()) cubin = it->second;
    }
    if (cubin.empty()) {
    nvrtcProgram np;
    if (a.createProgram(&np, src.c_str(), "sb_h_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { if (why) *why = "nvrtcCreateProgram failed"; return nullptr; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device"};
    nvrtcResult rc = a.compileProgram(np, 3, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t ls = 0;
        a.getLogSize(np, &ls);
        std::string log(ls, '\0');
        if (ls) a.getLog(np, &log[0]);
        if (why) *why = "NVRTC compile failed: " + log.substr(0, 400);
        a.destroyProgram(&np);
        return nullptr;
    }
    size_t cs = 0;
    a.getCubinSize(np, &cs);
    cubin.assign(cs, '\0');
    a.getCubin(np, &cubin[0]);
    a.destroyProgram(&np);
    std::lock_guard<std::mutex> lk(cache_mu);
    cubin_cache[src] = cubin;
    }
    ExprJit *j = new ExprJit();
    if (a.moduleLoadData(&j->mod, cubin.data()) != CUDA_SUCCESS || a.moduleGetFunction(&j->fn, j->mod, "sb_h_jit") != CUDA_SUCCESS) {
        if (why) *why = "cuModuleLoadData / cuModuleGetFunction failed";
        expr_jit_free(j);
        return nullptr;
    }
    j->n_consts = (uint32_t)prog.consts.size();
    j->max_col = max_col;
    return j;
}

// out[i] = program(columns at row i) with the compiled kernel; `d_consts`: the program's constant table on the device (n_consts x 32 B)
int32_t expr_jit_run(sb_ctx *ctx, const ExprJit *j, const void *d_consts, const std::vector<const void *> &cols, uint32_t log_n, uint32_t rot_scale_log, void *d_out,
                     cudaStream_t st, const std::vector<uint8_t> *col_shift) {
    SB_REQUIRE(j && j->fn, "expr_jit_run: no kernel");
    SB_REQUIRE(cols.size() <= (size_t)JIT_MAX_COLS && j->max_col < cols.size(), "expr_jit_run: column table too small / too large");
    const uint64_t n = 1ull << log_n;
    SB_REQUIRE(n >= (uint64_t)JIT_THREADS, "expr_jit_run: domain smaller than one CTA");
    JitArgs A;
    memset(&A, 0, sizeof A);
    for (size_t c = 0; c < cols.size(); c++) {
        A.cols[c] = cols[c];
        A.shift[c] = col_shift ? (*col_shift)[c] : 0;
    }
    A.consts = d_consts;
    A.out = d_out;
    A.mask = n - 1;
    A.rot_mul = 1ll << rot_scale_log;
    void *params[] = {&A};
    CUresult r = api().launchKernel(j->fn, (unsigned)(n / JIT_THREADS), 1, 1, JIT_THREADS, 1, 1, 0, (CUstream)st, params, nullptr);
    if (r != CUDA_SUCCESS) {
        set_last_error("expr_jit_run: cuLaunchKernel failed (%d)", (int)r);
        return SB_ERR_CUDA;
    }
    ctx->launches++;
    return SB_OK;
}

}  // namespace sb
