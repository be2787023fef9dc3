// Create-time kernel baking for FSSPMDM.  Default: the kernel is emitted as PTX text and assembled by the
// driver (cudaLibraryLoadData).  Alternative (LIBXSMM_B200_FSSPMDM_JIT=nvrtc): CUDA C++ compiled by NVRTC, loaded
// lazily with dlopen so that the library itself loads on machines without a CUDA toolkit.  On any failure
// create() falls back to the generic kernel and records the reason.
#include "fsspmdm_jit.h"
#include "common.cuh"
#include <nvrtc.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdint>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <mutex>

namespace xb {

struct FsJit {
  cudaLibrary_t lib;
  cudaKernel_t kern;
  int cols_per_thread;  // 1, or 2 (float, even pitches: 8-byte accesses)
  int block;
};

namespace {

struct Nvrtc {
  void* h;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
};

Nvrtc* nvrtc()
{
  static Nvrtc api;
  static int state = 0;   // 0 unknown, 1 ok, -1 unavailable
  static std::mutex mtx;
  std::lock_guard<std::mutex> lock(mtx);
  if (0 == state) {
    const char* names[] = { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so" };
    api.h = 0;
    for (size_t i = 0; i < sizeof(names) / sizeof(*names) && 0 == api.h; ++i) api.h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    state = -1;
    if (api.h) {
      *(void**)&api.CreateProgram = dlsym(api.h, "nvrtcCreateProgram");
      *(void**)&api.CompileProgram = dlsym(api.h, "nvrtcCompileProgram");
      *(void**)&api.GetCUBINSize = dlsym(api.h, "nvrtcGetCUBINSize");
      *(void**)&api.GetCUBIN = dlsym(api.h, "nvrtcGetCUBIN");
      *(void**)&api.GetProgramLogSize = dlsym(api.h, "nvrtcGetProgramLogSize");
      *(void**)&api.GetProgramLog = dlsym(api.h, "nvrtcGetProgramLog");
      *(void**)&api.DestroyProgram = dlsym(api.h, "nvrtcDestroyProgram");
      if (api.CreateProgram && api.CompileProgram && api.GetCUBINSize && api.GetCUBIN && api.DestroyProgram) state = 1;
    }
  }
  return 1 == state ? &api : 0;
}

const int kBlock = 128;

void append(std::string& s, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  s += buf;
}

// The emitted kernel: one thread per CPT consecutive columns of B/C (CPT = 1, or 2 as a vector of two).
// Every B row that the operator touches is loaded once into a named register (coalesced across the warp),
// every output row is then an in-order chain of fused multiply-adds with literal operator values -- the
// rounding sequence of the reference's emitted kernel (generator :227-300) -- and is stored, with a
// streaming store, as soon as it is complete.
void emit_kernel(std::string& s, const char* name, int cpt, int is_double, int M, int K, int beta_one, int skip_empty,
                 const int* rowptr, const int* col, const double* val)
{
  // Only scalar variables are emitted (the two columns of the vector form are split right after the load and
  // joined right before the store): NVRTC's optimiser needs several times longer for thousands of float2 values.
  const char* T = is_double ? "double" : "float";
  char V[16];
  snprintf(V, sizeof(V), "%s%s", T, 2 == cpt ? "2" : "");
  const char* f = is_double ? "fma" : "fmaf";
  std::vector<char> used(K, 0);
  for (int u = 0; u < rowptr[M]; ++u) used[col[u]] = 1;
  append(s, "extern \"C\" __global__ void __launch_bounds__(%d) %s(const %s* __restrict__ B, %s* __restrict__ C, long long ncols, long long ldb, long long ldc)\n{\n", kBlock, name, T, T);
  append(s, "  const long long n = ((long long)blockIdx.x * %d + threadIdx.x) * %d;\n  if (n >= ncols) return;\n", kBlock, cpt);
  append(s, "  const %s* __restrict__ b = B + n;\n  %s* __restrict__ c = C + n;\n  %s ax%s;\n", T, T, T, 2 == cpt ? ", ay" : "");
  for (int k = 0; k < K; ++k) if (used[k]) {
    if (2 == cpt) append(s, "  const %s t%d = __ldcs((const %s*)(b + %dLL * ldb)); const %s b%dx = t%d.x, b%dy = t%d.y;\n", V, k, V, k, T, k, k, k, k);
    else append(s, "  const %s b%dx = __ldcs(b + %dLL * ldb);\n", T, k, k);
  }
  // beta = 1: the C rows of a group are fetched together ahead of the group's chains, so that their latency
  // is paid once per group instead of once per row (loads cannot be hoisted over the stores by the compiler)
  const int group = beta_one ? (is_double ? 8 : 16) : 1;
  for (int m0 = 0; m0 < M; m0 += group) {
    if (beta_one) {
      for (int m = m0; m < M && m < m0 + group; ++m) if (rowptr[m + 1] != rowptr[m]) {
        if (2 == cpt) append(s, "  const %s u%d = __ldcs((const %s*)(c + %dLL * ldc)); const %s c%dx = u%d.x, c%dy = u%d.y;\n", V, m, V, m, T, m, m, m, m);
        else append(s, "  const %s c%dx = __ldcs(c + %dLL * ldc);\n", T, m, m);
      }
    }
    for (int m = m0; m < M && m < m0 + group; ++m) {
      const int lo = rowptr[m], hi = rowptr[m + 1];
      if (hi == lo) {
        if (!skip_empty && !beta_one) {
          if (2 == cpt) append(s, "  __stcs((%s*)(c + %dLL * ldc), make_%s(0, 0));\n", V, m, V);
          else append(s, "  __stcs(c + %dLL * ldc, (%s)0);\n", m, T);
        }
        continue;
      }
      if (beta_one) append(s, 2 == cpt ? "  ax = c%dx; ay = c%dy;\n" : "  ax = c%dx;\n", m, m);
      else append(s, 2 == cpt ? "  ax = 0; ay = 0;\n" : "  ax = 0;\n");
      for (int u = lo; u < hi; ++u) {
        char lit[64];
        if (is_double) {
          long long bits; const double v = val[u];
          memcpy(&bits, &v, 8);
          snprintf(lit, sizeof(lit), "__longlong_as_double(0x%016llxLL)", (unsigned long long)bits);
        }
        else {
          int bits; const float v = (float)val[u];
          memcpy(&bits, &v, 4);
          snprintf(lit, sizeof(lit), "__int_as_float(0x%08x)", (unsigned)bits);
        }
        if (2 == cpt) append(s, "  ax = %s(%s, b%dx, ax); ay = %s(%s, b%dy, ay);\n", f, lit, col[u], f, lit, col[u]);
        else append(s, "  ax = %s(%s, b%dx, ax);\n", f, lit, col[u]);
      }
      if (2 == cpt) append(s, "  __stcs((%s*)(c + %dLL * ldc), make_%s(ax, ay));\n", V, m, V);
      else append(s, "  __stcs(c + %dLL * ldc, ax);\n", m);
    }
  }
  s += "}\n";
}

// vec2: emit the two-columns-per-thread form (8-byte accesses; float only) instead of the one-column form
std::string emit(int is_double, int vec2, int M, int K, int beta_one, int skip_empty,
                 const int* rowptr, const int* col, const double* val)
{
  std::string s;
  s.reserve(160 * (size_t)rowptr[M] + 8192);
  emit_kernel(s, "fs_baked", (vec2 && !is_double) ? 2 : 1, is_double, M, K, beta_one, skip_empty, rowptr, col, val);
  return s;
}

// The same kernel as PTX text.  The driver's JIT (ptxas) assembles thousands of straight-line fused
// multiply-adds in a fraction of the time NVRTC's optimiser needs for the CUDA C++ form (measured: 6.3 s vs
// well under a second for the 150 x 64 / 30 % float operator), which is what keeps create() interactive.
static int fs_prefetch_b()
{
  const char* e = getenv("LIBXSMM_B200_FSSPMDM_PREFETCH_B");   // developer switch: distance of the B prefetch in half waves (default 2 = one wave; 0 = off)
  return (e && *e) ? atoi(e) : 2;
}

static int fs_prefetch_groups()
{
  const char* e = getenv("LIBXSMM_B200_FSSPMDM_PREFETCH");     // developer switch: groups of C rows prefetched ahead (0 = off)
  return (e && *e) ? atoi(e) : 2;
}

std::string emit_ptx(int is_double, int vec2, int M, int K, int beta_one, int skip_empty,
                     const int* rowptr, const int* col, const double* val)
{
  const int cpt = (vec2 && !is_double) ? 2 : 1;
  const int esz = is_double ? 8 : 4;
  const char* ty = is_double ? "f64" : "f32";
  std::vector<char> used(K, 0);
  for (int u = 0; u < rowptr[M]; ++u) used[col[u]] = 1;
  std::string s;
  s.reserve(96 * (size_t)rowptr[M] * cpt + 8192);
  s += ".version 8.6\n.target sm_100a\n.address_size 64\n\n";
  s += ".visible .entry fs_baked(.param .u64 pB, .param .u64 pC, .param .u64 pN, .param .u64 pLDB, .param .u64 pLDC)\n";
  append(s, ".maxntid %d, 1, 1\n", kBlock);
  s += "{\n";
  s += "  .reg .pred %p, %pf;\n  .reg .b32 %r<4>;\n  .reg .b64 %rd<12>;\n";
  append(s, "  .reg .%s %%bx<%d>, %%by<%d>, %%cx<%d>, %%cy<%d>, %%ax, %%ay;\n", ty, K, K, M, M);
  s += "  ld.param.u64 %rd0, [pB];\n  ld.param.u64 %rd1, [pC];\n  ld.param.u64 %rd2, [pN];\n  ld.param.u64 %rd3, [pLDB];\n  ld.param.u64 %rd4, [pLDC];\n";
  s += "  mov.u32 %r0, %ctaid.x;\n  mov.u32 %r1, %tid.x;\n";
  append(s, "  mul.wide.u32 %%rd5, %%r0, %d;\n  cvt.u64.u32 %%rd6, %%r1;\n  add.s64 %%rd5, %%rd5, %%rd6;\n", kBlock);
  if (2 == cpt) s += "  shl.b64 %rd5, %rd5, 1;\n";
  s += "  setp.ge.s64 %p, %rd5, %rd2;\n  @%p bra DONE;\n";
  append(s, "  mad.lo.s64 %%rd0, %%rd5, %d, %%rd0;\n  mad.lo.s64 %%rd1, %%rd5, %d, %%rd1;\n", esz, esz);   // b = B + n, c = C + n
  append(s, "  mul.lo.s64 %%rd3, %%rd3, %d;\n  mul.lo.s64 %%rd4, %%rd4, %d;\n", esz, esz);                 // row pitches in bytes
  s += "  and.b32 %r2, %r1, 15;\n  setp.eq.u32 %pf, %r2, 0;\n";                                          // one lane per 128-byte line issues prefetches
  for (int k = 0; k < K; ++k) if (used[k]) {
    append(s, "  mad.lo.s64 %%rd7, %%rd3, %d, %%rd0;\n", k);
    if (2 == cpt) append(s, "  ld.global.cs.v2.%s {%%bx%d, %%by%d}, [%%rd7];\n", ty, k, k);
    else append(s, "  ld.global.cs.%s %%bx%d, [%%rd7];\n", ty, k);
  }
  // The B rows of the column block one "wave" ahead (number of SMs x this block's columns further on: blocks are
  // scheduled in order) are pulled into L2 now, one lane per 128-byte line: that block's load phase then pays L2
  // latency instead of DRAM latency.  Measured on B200 (fraction of the 6554 GB/s copy peak): 150 x 64 float operator,
  // N = 2^24: 0.854 -> 0.989; double, N = 2^20: 0.960 -> 0.984.  Twice as far ahead is worth nothing (0.88), further
  // is harmful; with beta = 1 the C-row prefetches below already fill the queues and this one costs 1-4 %.
  if (!beta_one && fs_prefetch_b() > 0) {      // (measured: helps beta = 0, costs 2-4 % with beta = 1, where the C prefetches below already fill the queues)
    s += "  mov.u32 %r3, %nsmid;\n";
    append(s, "  mul.wide.u32 %%rd10, %%r3, %d;\n", fs_prefetch_b() * kBlock * cpt / 2);                                          // columns ahead
    s += "  add.s64 %rd11, %rd5, %rd10;\n  setp.lt.s64 %p, %rd11, %rd2;\n  and.pred %p, %p, %pf;\n";
    append(s, "  mad.lo.s64 %%rd10, %%rd10, %d, %%rd0;\n", esz);
    for (int k = 0; k < K; ++k) if (used[k]) append(s, "  mad.lo.s64 %%rd7, %%rd3, %d, %%rd10;\n  @%%p prefetch.global.L2 [%%rd7];\n", k);
  }
  const int group = beta_one ? (is_double ? 8 : 16) : 1;
  // beta = 1: the C rows that will be read are pulled into L2 two groups ahead by prefetch instructions (one lane
  // per 128-byte line, no registers held), so that the group's loads pay the L2 latency instead of the DRAM one
  const int ahead = fs_prefetch_groups();
  auto prefetch_group = [&](int g0) {
    for (int m = g0; m < M && m < g0 + group; ++m) if (rowptr[m + 1] != rowptr[m]) {
      append(s, "  mad.lo.s64 %%rd9, %%rd4, %d, %%rd1;\n  @%%pf prefetch.global.L2 [%%rd9];\n", m);
    }
  };
  if (beta_one && ahead > 0) for (int a = 0; a < ahead; ++a) prefetch_group(a * group);
  for (int m0 = 0; m0 < M; m0 += group) {
    if (beta_one) {
      if (ahead > 0) prefetch_group(m0 + ahead * group);
      for (int m = m0; m < M && m < m0 + group; ++m) if (rowptr[m + 1] != rowptr[m]) {
        append(s, "  mad.lo.s64 %%rd8, %%rd4, %d, %%rd1;\n", m);
        if (2 == cpt) append(s, "  ld.global.cs.v2.%s {%%cx%d, %%cy%d}, [%%rd8];\n", ty, m, m);
        else append(s, "  ld.global.cs.%s %%cx%d, [%%rd8];\n", ty, m);
      }
    }
    for (int m = m0; m < M && m < m0 + group; ++m) {
      const int lo = rowptr[m], hi = rowptr[m + 1];
      const char* zero = is_double ? "0d0000000000000000" : "0f00000000";
      if (hi == lo) {
        if (!skip_empty && !beta_one) {
          append(s, "  mad.lo.s64 %%rd8, %%rd4, %d, %%rd1;\n  mov.%s %%ax, %s;\n", m, ty, zero);
          if (2 == cpt) append(s, "  st.global.cs.v2.%s [%%rd8], {%%ax, %%ax};\n", ty);
          else append(s, "  st.global.cs.%s [%%rd8], %%ax;\n", ty);
        }
        continue;
      }
      if (beta_one) {
        append(s, "  mov.%s %%ax, %%cx%d;\n", ty, m);
        if (2 == cpt) append(s, "  mov.%s %%ay, %%cy%d;\n", ty, m);
      }
      else {
        append(s, "  mov.%s %%ax, %s;\n", ty, zero);
        if (2 == cpt) append(s, "  mov.%s %%ay, %s;\n", ty, zero);
      }
      for (int u = lo; u < hi; ++u) {
        char lit[32];
        if (is_double) {
          unsigned long long bits; const double v = val[u];
          memcpy(&bits, &v, 8);
          snprintf(lit, sizeof(lit), "0d%016llX", bits);
        }
        else {
          unsigned int bits; const float v = (float)val[u];
          memcpy(&bits, &v, 4);
          snprintf(lit, sizeof(lit), "0f%08X", bits);
        }
        append(s, "  fma.rn.%s %%ax, %s, %%bx%d, %%ax;\n", ty, lit, col[u]);
        if (2 == cpt) append(s, "  fma.rn.%s %%ay, %s, %%by%d, %%ay;\n", ty, lit, col[u]);
      }
      append(s, "  mad.lo.s64 %%rd8, %%rd4, %d, %%rd1;\n", m);
      if (2 == cpt) append(s, "  st.global.cs.v2.%s [%%rd8], {%%ax, %%ay};\n", ty);
      else append(s, "  st.global.cs.%s [%%rd8], %%ax;\n", ty);
    }
  }
  s += "DONE:\n  ret;\n}\n";
  return s;
}

bool supported(int is_double, int M, int K, const int* rowptr, const int* col)
{
  std::vector<char> used(K, 0);
  int nused = 0;
  for (int u = 0; u < rowptr[M]; ++u) if (!used[col[u]]) { used[col[u]] = 1; ++nused; }
  // B rows live in registers: 2 registers per double, 1 per float, out of 255
  // B rows live in registers (2 registers per double, 1 per float, out of 255); the NVRTC form is additionally
  // limited to ~6000 nonzeros in fs_jit_build (compile time grows faster than linearly)
  return rowptr[M] > 0 && (is_double ? nused <= 100 : nused <= 200) && rowptr[M] <= 40000;
}

std::mutex g_cache_mtx;
std::map<std::string, std::vector<char> > g_cubin_cache;

}  // namespace

char* fs_jit_source(int is_double, int vec2, int M, int K, int beta_one, int skip_empty,
                    const int* rowptr, const int* col, const double* val)
{
  const char* env = getenv("LIBXSMM_B200_FSSPMDM_JIT");
  const std::string s = (env && 'n' == *env) ? emit(is_double, vec2, M, K, beta_one, skip_empty, rowptr, col, val)
                                             : emit_ptx(is_double, vec2, M, K, beta_one, skip_empty, rowptr, col, val);
  char* out = (char*)malloc(s.size() + 1);
  if (out) memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

FsJit* fs_jit_build(int is_double, int vec2, int M, int K, int beta_one, int skip_empty,
                    const int* rowptr, const int* col, const double* val)
{
  const char* env = getenv("LIBXSMM_B200_FSSPMDM_JIT");
  if (env && '0' == *env) return 0;
  if (!supported(is_double, M, K, rowptr, col)) return 0;
  // default: PTX text assembled by the driver (fast); LIBXSMM_B200_FSSPMDM_JIT=nvrtc: CUDA C++ through NVRTC
  const bool use_nvrtc = (env && 'n' == *env);
  std::vector<char> cubin;     // cubin (NVRTC) or NUL-terminated PTX text
  if (!use_nvrtc) {
    const std::string ptx = emit_ptx(is_double, vec2, M, K, beta_one, skip_empty, rowptr, col, val);
    cubin.assign(ptx.begin(), ptx.end());
    cubin.push_back(0);
  }
  else {
    if (rowptr[M] > 6000) return 0;
    Nvrtc* rt = nvrtc();
    if (0 == rt) { set_error(-2, "fsspmdm: NVRTC (libnvrtc.so.12) not found; using the generic kernel"); return 0; }
    const std::string src = emit(is_double, vec2, M, K, beta_one, skip_empty, rowptr, col, val);
    {
      std::lock_guard<std::mutex> lock(g_cache_mtx);
      std::map<std::string, std::vector<char> >::const_iterator it = g_cubin_cache.find(src);
      if (it != g_cubin_cache.end()) cubin = it->second;
    }
    if (cubin.empty()) {
      nvrtcProgram prog = 0;
      if (NVRTC_SUCCESS != rt->CreateProgram(&prog, src.c_str(), "fs_baked.cu", 0, 0, 0)) { set_error(-3, "nvrtcCreateProgram failed"); return 0; }
      const char* opts[] = { "--gpu-architecture=sm_100a", "-lineinfo", "--fmad=false" };
      const nvrtcResult rc = rt->CompileProgram(prog, 3, opts);
      if (NVRTC_SUCCESS != rc) {
        size_t n = 0;
        std::string log;
        if (rt->GetProgramLogSize && NVRTC_SUCCESS == rt->GetProgramLogSize(prog, &n) && n > 1) { log.resize(n); rt->GetProgramLog(prog, &log[0]); }
        set_error(-4, "fsspmdm: NVRTC compile failed (%d): %.300s", (int)rc, log.c_str());
        rt->DestroyProgram(&prog);
        return 0;
      }
      size_t sz = 0;
      if (NVRTC_SUCCESS != rt->GetCUBINSize(prog, &sz) || 0 == sz) { set_error(-5, "nvrtcGetCUBINSize failed"); rt->DestroyProgram(&prog); return 0; }
      cubin.resize(sz);
      rt->GetCUBIN(prog, cubin.data());
      rt->DestroyProgram(&prog);
      std::lock_guard<std::mutex> lock(g_cache_mtx);
      g_cubin_cache[src] = cubin;
    }
  }
  FsJit* j = new FsJit();
  j->block = kBlock;
  j->cols_per_thread = (vec2 && !is_double) ? 2 : 1;
  cudaError_t e = cudaLibraryLoadData(&j->lib, cubin.data(), 0, 0, 0, 0, 0, 0);
  if (cudaSuccess == e) e = cudaLibraryGetKernel(&j->kern, j->lib, "fs_baked");
  if (cudaSuccess != e) {
    set_error((int)e, "fsspmdm: loading the baked kernel failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    delete j;
    return 0;
  }
  return j;
}

bool fs_jit_launch(const FsJit* j, const void* dB, void* dC, long long ncols, long long ldb, long long ldc, cudaStream_t stream)
{
  // the 2-column form needs 8-byte aligned rows: even pitches, even column count, aligned bases; a panel that does
  // not qualify goes to the generic kernel (return false)
  if (2 == j->cols_per_thread && !((0 == ((ldb | ldc | ncols) & 1)) && (0 == (((uintptr_t)dB | (uintptr_t)dC) & 7)))) return false;
  const long long per_block = (long long)j->block * j->cols_per_thread;
  const long long blocks = (ncols + per_block - 1) / per_block;
  void* args[] = { (void*)&dB, (void*)&dC, (void*)&ncols, (void*)&ldb, (void*)&ldc };
  const cudaError_t e = cudaLaunchKernel((const void*)j->kern, dim3((unsigned)blocks, 1, 1), dim3((unsigned)j->block, 1, 1), args, 0, stream);
  if (cudaSuccess != e) { set_error((int)e, "fsspmdm: baked kernel launch failed: %s", cudaGetErrorString(e)); return false; }
  return true;
}

void fs_jit_destroy(FsJit* j)
{
  if (0 == j) return;
  cudaLibraryUnload(j->lib);
  delete j;
}

}  // namespace xb
