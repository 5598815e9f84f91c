// Plan object: cuFFT plans + partition of the caller's workspace.
#include "common.cuh"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace jps {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- launch accounting
static const char* kKernelNames[K_NUM] = {
    "paint_atomic", "bucket_count", "bucket_scan", "bucket_scatter", "bucket_fine", "paint_tile", "pk_fold_bin",
    "pk_count_modes", "pk_finalize", "cufft_r2c", "cufft_c2r", "memset", "shell_filter",
    "triple_reduce", "xi_bin", "misc", "text_index", "text_parse", "text_compact",
    "mock_field", "mock_populate", "interlace_combine", "fft_transpose", "cufft_c2c_y", "cufft_c2c_x"};

// Distinct plans may be driven from distinct host threads (include/jps.h), so the process-wide
// accounting is atomic counters + one mutex around the event lists.
struct Pending { int id; cudaEvent_t e0, e1; };
static std::atomic<bool> g_prof_on{false};
static std::atomic<unsigned long long> g_launches[K_NUM];
static double g_ms[K_NUM];                        // guarded by g_prof_mutex
static std::vector<Pending> g_pending;            // guarded by g_prof_mutex
static std::vector<cudaEvent_t> g_pool;           // guarded by g_prof_mutex
static std::mutex g_prof_mutex;

static cudaEvent_t get_event() {
  {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ScopedLaunch::ScopedLaunch(int id_, cudaStream_t s_) : id(id_), s(s_), e0(nullptr), e1(nullptr), timed(g_prof_on.load()) {
  g_launches[id].fetch_add(1, std::memory_order_relaxed);
  if (timed) { e0 = get_event(); e1 = get_event(); cudaEventRecord(e0, s); }
}

ScopedLaunch::~ScopedLaunch() {
  if (timed) {
    cudaEventRecord(e1, s);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_pending.push_back({id, e0, e1});
  }
}

static void drain_pending() {
  std::vector<Pending> todo;
  {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    todo.swap(g_pending);
  }
  for (auto& p : todo) {                          // synchronise outside the lock
    float ms = 0.f;
    const bool ok = cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (ok) g_ms[p.id] += ms;
    g_pool.push_back(p.e0); g_pool.push_back(p.e1);
  }
}

// (1/sinc(pi k/N))^p per axis in float32, operation by operation as
// /root/reference/src/correlations.py:15,20-21,32 evaluates it (Q13):
//   prefact = pi/dims (Python double, cast when it meets the int32 k vector)
//   x = prefact*k ; y = x/pi ; sinc(y) = sin(pi*y)/(pi*y), 1 at y == 0 ; (1/sinc)**p
void host_window_axis(int n, int p, float* out) {
  const float pref = (float)(M_PI / (double)n);
  const float pi32 = (float)M_PI;
  const int mid = n / 2;
  for (int i = 0; i < n; ++i) {
    const int ki = i > mid ? i - n : i;
    const float x = pref * (float)ki;
    const float y = x / pi32;
    float s = 1.0f;
    if (y != 0.0f) {
      const float pix = pi32 * y;
      s = sinf(pix) / pix;
    }
    const float r = 1.0f / s;
    float v = r;
    for (int j = 1; j < p; ++j) v = v * r;
    out[i] = v;
  }
}


struct TableLayout {
  size_t lut, compact_to_bin, bin_to_compact, edges, cnt, ksum, lastidx, seg_bp, seg_val, coarse;
};

struct Layout {
  size_t dk, dk2, ztw, fft_work, wlut, acc, scal, isum, shell, total;
  TableLayout t[kNumTables];
  int cap;
};

static Layout make_layout(int n, int pitch, size_t fft_work_bytes, int n_shell_fields, bool tables_only = false,
                          bool pencil = false) {
  Layout L;
  const int mid = n / 2;
  const int64_t k2max = 3LL * mid * mid;
  L.cap = (int)std::min<int64_t>(k2max + 2, 262144);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.dk = take(tables_only ? 0 : (size_t)n * n * pitch * sizeof(float2));
  L.dk2 = take(pencil ? (size_t)n * n * pitch * sizeof(float2) : 0);
  L.ztw = take(pencil ? (size_t)(n / 4 + 2) * sizeof(float2) : 0);
  L.fft_work = take(fft_work_bytes);
  for (int i = 0; i < kNumTables; ++i) {
    L.t[i].lut = take((size_t)(k2max + 1) * 4);
    L.t[i].compact_to_bin = take((size_t)L.cap * 4);
    L.t[i].bin_to_compact = take((size_t)kMaxUserBins * 4);
    L.t[i].edges = take((size_t)(kMaxUserBins + 1) * 4);
    L.t[i].cnt = take((size_t)L.cap * 8);
    L.t[i].ksum = take((size_t)L.cap * 8);
    L.t[i].lastidx = take((size_t)L.cap * 8);
    L.t[i].seg_bp = take((size_t)kMaxSegments * 4);
    L.t[i].seg_val = take((size_t)kMaxSegments * 4);
    L.t[i].coarse = take((size_t)((int64_t)sqrt((double)k2max) + 4) * 4);
  }
  L.wlut = take((size_t)3 * n * 4);
  L.acc = take((size_t)L.cap * 4 * 8);
  L.scal = take((size_t)1024 * 8);
  L.isum = take((size_t)kIsumSlots * 8);
  L.shell = take((size_t)n_shell_fields * n * n * pitch * sizeof(float2));
  L.total = off;
  return L;
}

static int make_r2c(int n, int pitch, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  long long dims[3] = {n, n, n};
  long long inembed[3] = {n, n, n};
  long long onembed[3] = {n, n, pitch};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 3, dims, inembed, 1, (long long)n * n * n, onembed, 1,
                                      (long long)n * n * pitch, CUFFT_R2C, 1, work));
  return JPS_OK;
}

static int make_c2r_inplace(int n, int pitch, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  long long dims[3] = {n, n, n};
  long long inembed[3] = {n, n, pitch};
  long long onembed[3] = {n, n, 2LL * pitch};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 3, dims, inembed, 1, (long long)n * n * pitch, onembed, 1,
                                      2LL * n * n * pitch, CUFFT_C2R, 1, work));
  return JPS_OK;
}

static int make_r2c_inplace(int n, int pitch, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  long long dims[3] = {n, n, n};
  long long inembed[3] = {n, n, 2LL * pitch};
  long long onembed[3] = {n, n, pitch};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 3, dims, inembed, 1, 2LL * n * n * pitch, onembed, 1,
                                      (long long)n * n * pitch, CUFFT_R2C, 1, work));
  return JPS_OK;
}

// contiguous batched 1-D plans of the pencil decomposition

static int make_pencil_c2c(int n, long long batch, cufftHandle* h, size_t* work) {   // C2C, `batch` lines of n
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  long long dims[1] = {n};
  long long embed[1] = {n};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 1, dims, embed, 1, n, embed, 1, n, CUFFT_C2C, batch, work));
  return JPS_OK;
}

static int pitch_for(int n) { return n / 2 + 1; }


}  // namespace jps

using namespace jps;

extern "C" int jps_profile_enable(int on) {
  drain_pending();
  g_prof_on.store(on != 0);
  return JPS_OK;
}

extern "C" int jps_profile_reset(void) {
  drain_pending();
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (int i = 0; i < K_NUM; ++i) { g_launches[i].store(0); g_ms[i] = 0.0; }
  return JPS_OK;
}

extern "C" int jps_profile_num_kernels(void) { return K_NUM; }

extern "C" int jps_profile_get(int id, const char** name, unsigned long long* launches, double* ms) {
  JPS_REQUIRE(id >= 0 && id < K_NUM, "jps_profile_get: id out of range");
  drain_pending();                       // synchronises on the recorded events
  if (name) *name = kKernelNames[id];
  if (launches) *launches = g_launches[id].load();
  if (ms) { std::lock_guard<std::mutex> lock(g_prof_mutex); *ms = g_ms[id]; }
  return JPS_OK;
}

extern "C" int jps_version(void) { return JPS_VERSION; }
extern "C" const char* jps_last_error(void) { return g_err; }

extern "C" int jps_plan_workspace_bytes(int n_mesh, int n_shell_fields, int flags, size_t* bytes) {
  JPS_REQUIRE(bytes != nullptr, "jps_plan_workspace_bytes: bytes is NULL");
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_plan_workspace_bytes: n_mesh=%d out of range [2,4096]", n_mesh);
  JPS_REQUIRE(n_shell_fields >= 0, "jps_plan_workspace_bytes: n_shell_fields < 0");
  const int pitch = pitch_for(n_mesh);
  if (flags & JPS_PLAN_TABLES_ONLY) {
    *bytes = make_layout(n_mesh, pitch, 0, 0, true).total;
    return JPS_OK;
  }
  cufftHandle h;
  size_t w1 = 0, w2 = 0;
  if (flags & JPS_PLAN_FFT_PENCIL) {
    JPS_REQUIRE(n_shell_fields == 0, "jps_plan_workspace_bytes: a JPS_PLAN_FFT_PENCIL plan has no shell fields");
    size_t wz = 0, wy = 0;
    JPS_REQUIRE(n_mesh % 2 == 0, "jps_plan_workspace_bytes: JPS_PLAN_FFT_PENCIL needs an even mesh size");
    int rcp = make_pencil_c2c(n_mesh / 2, (long long)n_mesh * n_mesh, &h, &wz);
    cufftDestroy(h);
    if (rcp) return rcp;
    rcp = make_pencil_c2c(n_mesh, (long long)n_mesh * pitch, &h, &wy);
    cufftDestroy(h);
    if (rcp) return rcp;
    *bytes = make_layout(n_mesh, pitch, std::max(wz, wy), 0, false, true).total;
    return JPS_OK;
  }
  int rc = make_r2c(n_mesh, pitch, &h, &w1);
  cufftDestroy(h);
  if (rc) return rc;
  rc = make_c2r_inplace(n_mesh, pitch, &h, &w2);
  cufftDestroy(h);
  if (rc) return rc;
  size_t w3 = 0;
  if (n_shell_fields > 0) {
    rc = make_r2c_inplace(n_mesh, pitch, &h, &w3);
    cufftDestroy(h);
    if (rc) return rc;
  }
  *bytes = make_layout(n_mesh, pitch, std::max(std::max(w1, w2), w3), n_shell_fields).total;
  return JPS_OK;
}

extern "C" int jps_plan_create(int n_mesh, int n_shell_fields, int flags, void* workspace,
                               size_t workspace_bytes, jps_plan_t** out) {
  const bool tables_only = (flags & JPS_PLAN_TABLES_ONLY) != 0;
  JPS_REQUIRE(out != nullptr, "jps_plan_create: plan is NULL");
  *out = nullptr;
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_plan_create: n_mesh=%d out of range [2,4096]", n_mesh);
  JPS_REQUIRE(workspace != nullptr, "jps_plan_create: workspace is NULL");
  JPS_REQUIRE(((uintptr_t)workspace & 255) == 0, "jps_plan_create: workspace must be 256-byte aligned");
  jps_plan* p = new jps_plan();
  p->n = n_mesh;
  p->nz = n_mesh / 2 + 1;
  p->pitch = pitch_for(n_mesh);
  p->n_shell_fields = n_shell_fields;
  p->k2max = 3LL * (n_mesh / 2) * (n_mesh / 2);
  cudaError_t ce = cudaGetDevice(&p->device);
  if (ce != cudaSuccess) { delete p; set_error("jps_plan_create: cudaGetDevice failed: %s", cudaGetErrorString(ce)); return JPS_ERR_CUDA; }
  size_t w1 = 0, w2 = 0, w3 = 0;
  int rc = JPS_OK;
  const bool pencil = (flags & JPS_PLAN_FFT_PENCIL) != 0;
  if (pencil) {
    if (tables_only || n_shell_fields != 0) {
      delete p;
      set_error("jps_plan_create: JPS_PLAN_FFT_PENCIL excludes JPS_PLAN_TABLES_ONLY and shell fields");
      return JPS_ERR_INVALID;
    }
    if (n_mesh % 2) { delete p; set_error("jps_plan_create: JPS_PLAN_FFT_PENCIL needs an even mesh size"); return JPS_ERR_INVALID; }
    p->pencil = true;
    // z-pass: the n reals of a line are read as n/2 complex numbers (even, odd samples): C2C of length n/2, then the
    // untangle step of the real transform is done by OUR transposing kernel on the way (r2c_untangle_transpose)
    rc = make_pencil_c2c(n_mesh / 2, (long long)n_mesh * n_mesh, &p->fz, &w1);
    if (rc) { delete p; return rc; }
    p->fz_ok = true;
    rc = make_pencil_c2c(n_mesh, (long long)n_mesh * p->pitch, &p->fy, &w2);
    if (rc) { jps_plan_destroy(p); return rc; }
    p->fy_ok = true;
    rc = make_pencil_c2c(n_mesh, (long long)n_mesh * p->pitch, &p->fx, &w3);
    if (rc) { jps_plan_destroy(p); return rc; }
    p->fx_ok = true;
  } else if (!tables_only) {
    rc = make_r2c(n_mesh, p->pitch, &p->r2c, &w1);
    if (rc) { delete p; return rc; }
    p->r2c_ok = true;
    rc = make_c2r_inplace(n_mesh, p->pitch, &p->c2r, &w2);
    if (rc) { cufftDestroy(p->r2c); delete p; return rc; }
    p->c2r_ok = true;
    if (n_shell_fields > 0) {
      rc = make_r2c_inplace(n_mesh, p->pitch, &p->r2c_ip, &w3);
      if (rc) { cufftDestroy(p->r2c); cufftDestroy(p->c2r); delete p; return rc; }
      p->r2c_ip_ok = true;
    }
  }
  Layout L = make_layout(n_mesh, p->pitch, std::max(std::max(w1, w2), w3), n_shell_fields, tables_only, pencil);
  if (workspace_bytes < L.total) {
    set_error("jps_plan_create: workspace has %zu bytes, %zu needed", workspace_bytes, L.total);
    jps_plan_destroy(p);
    return JPS_ERR_WORKSPACE;
  }
  char* ws = (char*)workspace;
  p->ws = ws;
  p->ws_bytes = workspace_bytes;
  p->dk = (float2*)(ws + L.dk);
  p->dk2 = pencil ? (float2*)(ws + L.dk2) : nullptr;
  p->ztw = pencil ? (float2*)(ws + L.ztw) : nullptr;
  p->fft_work = ws + L.fft_work;
  p->fft_work_bytes = std::max(std::max(w1, w2), w3);
  for (int i = 0; i < kNumTables; ++i) {
    BinTable& T = p->tables[i];
    T.lut = (int32_t*)(ws + L.t[i].lut);
    T.compact_to_bin = (int32_t*)(ws + L.t[i].compact_to_bin);
    T.bin_to_compact = (int32_t*)(ws + L.t[i].bin_to_compact);
    T.edges = (float*)(ws + L.t[i].edges);
    T.cnt = (unsigned long long*)(ws + L.t[i].cnt);
    T.ksum = (double*)(ws + L.t[i].ksum);
    T.lastidx = (unsigned long long*)(ws + L.t[i].lastidx);
    T.seg_bp = (int32_t*)(ws + L.t[i].seg_bp);
    T.seg_val = (int32_t*)(ws + L.t[i].seg_val);
    T.coarse = (int32_t*)(ws + L.t[i].coarse);
  }
  p->wlut = (float*)(ws + L.wlut);
  p->acc = (double*)(ws + L.acc);
  p->scal = (double*)(ws + L.scal);
  p->isum = (double*)(ws + L.isum);
  p->shell = (float*)(ws + L.shell);
  p->acc_cap = L.cap;
  cufftResult r = CUFFT_SUCCESS;
  if (pencil) {
    r = cufftSetWorkArea(p->fz, p->fft_work);
    if (r == CUFFT_SUCCESS) r = cufftSetWorkArea(p->fy, p->fft_work);
    if (r == CUFFT_SUCCESS) r = cufftSetWorkArea(p->fx, p->fft_work);
  } else if (!tables_only) {
    r = cufftSetWorkArea(p->r2c, p->fft_work);
    if (r == CUFFT_SUCCESS) r = cufftSetWorkArea(p->c2r, p->fft_work);
    if (r == CUFFT_SUCCESS && p->r2c_ip_ok) r = cufftSetWorkArea(p->r2c_ip, p->fft_work);
  }
  if (r != CUFFT_SUCCESS) {
    set_error("jps_plan_create: cufftSetWorkArea failed (%d)", (int)r);
    jps_plan_destroy(p);
    return JPS_ERR_CUFFT;
  }
  if (pencil) {                                    // e^{-2 pi i k / n}, k = 0 .. n/4, evaluated in double
    std::vector<float2> tw;
    host_r2c_twiddles(n_mesh, tw);
    ce = cudaMemcpy(p->ztw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      set_error("jps_plan_create: twiddle table upload failed: %s", cudaGetErrorString(ce));
      jps_plan_destroy(p);
      return JPS_ERR_CUDA;
    }
  }
  // window tables for p = 2, 3, 4 (tiny; synchronous copy at plan creation)
  std::vector<float> w((size_t)3 * n_mesh);
  for (int q = 0; q < 3; ++q) host_window_axis(n_mesh, q + 2, w.data() + (size_t)q * n_mesh);
  ce = cudaMemcpy(p->wlut, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    set_error("jps_plan_create: window table upload failed: %s", cudaGetErrorString(ce));
    jps_plan_destroy(p);
    return JPS_ERR_CUDA;
  }
  *out = p;
  return JPS_OK;
}

extern "C" int jps_plan_destroy(jps_plan_t* p) {
  if (!p) return JPS_OK;
  if (p->r2c_ok) cufftDestroy(p->r2c);
  if (p->c2r_ok) cufftDestroy(p->c2r);
  if (p->r2c_ip_ok) cufftDestroy(p->r2c_ip);
  if (p->fz_ok) cufftDestroy(p->fz);
  if (p->fy_ok) cufftDestroy(p->fy);
  if (p->fx_ok) cufftDestroy(p->fx);
  delete p;
  return JPS_OK;
}
