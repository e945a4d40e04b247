// abi_core.cu -- library / context / problem upload entry points of include/ci_b200.h (no kernels).
#include "ci_host.cuh"
#include "ci_gibbs.cuh"

namespace {

using namespace ci;

template <typename R>
void build_tiles(const ci_problem* pb, const void* y_, const void* X_, int NB, int ld,
                 std::vector<R>& out) {
  const R* y = static_cast<const R*>(y_);
  const R* X = static_cast<const R*>(X_);
  const int T = pb->T, p = pb->p;
  const size_t te = (size_t)tile_elems(p);
  out.assign((size_t)NB * te, (R)0);
  const R qnan = std::numeric_limits<R>::quiet_NaN();
  for (int b = 0; b < NB; ++b) {
    R* tile = out.data() + (size_t)b * te;
    for (int tl = 0; tl < TB; ++tl) {
      const int t = b * TB + tl;
      R* row = tile + tile_off(tl, ld);
      if (t < T) {
        for (int j = 0; j < p; ++j) row[j] = X[(size_t)t * p + j];
        row[p] = y[t];
      } else {
        row[p] = qnan;   // padded step == masked step
      }
    }
  }
}

template <typename R>
int upload_batch(ci_ctx* c, const ci_problem* probs, int N, const void* y_, const void* X_,
                 const void* Om_) {
  const int T = probs[0].T, p = probs[0].p;
  const size_t te = (size_t)ci::tile_elems(p);
  const size_t tile_stride = (size_t)c->NB * te;                        // elements per series
  const size_t om_stride = (((size_t)p * p * sizeof(R) + 15) & ~(size_t)15) / sizeof(R) + 16 / sizeof(R);
  std::vector<R> tiles(tile_stride * N), om(om_stride * N, (R)0), gram((size_t)p * p * N + 4),
      xty((size_t)(p > 0 ? p : 1) * N + 4);
  c->b_prob.assign(probs, probs + N);
  c->b_yty.assign(N, 0.0); c->b_nobs.assign(N, 0);
  const R* y = static_cast<const R*>(y_);
  const R* X = static_cast<const R*>(X_);
  const R* Om = static_cast<const R*>(Om_);
  std::vector<R> one;
  for (int s = 0; s < N; ++s) {
    build_tiles<R>(&probs[s], y + (size_t)s * T, p ? X + (size_t)s * T * p : nullptr, c->NB, c->ld, one);
    std::copy(one.begin(), one.end(), tiles.begin() + tile_stride * s);
    if (p) std::copy(Om + (size_t)s * p * p, Om + (size_t)(s + 1) * p * p, om.begin() + om_stride * s);
    // sufficient statistics over observed rows, float64 on the host (as ci_set_data)
    std::vector<double> g((size_t)p * p, 0.0), b((size_t)(p > 0 ? p : 1), 0.0);
    double yty = 0.0; int nobs = 0;
    for (int t = 0; t < T; ++t) {
      const double yt = (double)y[(size_t)s * T + t];
      if (!(yt == yt)) continue;
      ++nobs; yty += yt * yt;
      const R* xr = X + ((size_t)s * T + t) * p;
      for (int i = 0; i < p; ++i) {
        b[i] += (double)xr[i] * yt;
        for (int j = 0; j <= i; ++j) g[(size_t)i * p + j] += (double)xr[i] * (double)xr[j];
      }
    }
    for (int i = 0; i < p; ++i)
      for (int j = i + 1; j < p; ++j) g[(size_t)i * p + j] = g[(size_t)j * p + i];
    for (int i = 0; i < p * p; ++i) gram[(size_t)s * p * p + i] = (R)g[i];
    for (int i = 0; i < p; ++i) xty[(size_t)s * p + i] = (R)b[i];
    c->b_yty[s] = yty; c->b_nobs[s] = nobs;
  }
  c->b_tile_stride = tile_stride * sizeof(R); c->b_omega_stride = om_stride * sizeof(R);
  c->b_gram_stride = (size_t)p * p * sizeof(R); c->b_xty_stride = (size_t)p * sizeof(R);
  CU_TRY(c->b_tiles.reserve(tiles.size() * sizeof(R)));
  CU_TRY(c->b_omega.reserve(om.size() * sizeof(R)));
  CU_TRY(c->b_gram.reserve(gram.size() * sizeof(R)));
  CU_TRY(c->b_xty.reserve(xty.size() * sizeof(R)));
  CU_TRY(cudaMemcpyAsync(c->b_tiles.p, tiles.data(), tiles.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_omega.p, om.data(), om.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_gram.p, gram.data(), gram.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_xty.p, xty.data(), xty.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  // per-series device descriptors of the batched kernels
  std::vector<ci::BatchDev<R>> dev(N);
  for (int s = 0; s < N; ++s) {
    c->prob = probs[s];
    c->v_tiles = static_cast<char*>(c->b_tiles.p) + c->b_tile_stride * s;
    c->v_omega = static_cast<char*>(c->b_omega.p) + c->b_omega_stride * s;
    dev[s].pr = make_probdev<R>(c);
    dev[s].gd.gram = reinterpret_cast<const R*>(static_cast<char*>(c->b_gram.p) + c->b_gram_stride * s);
    dev[s].gd.xty0 = reinterpret_cast<const R*>(static_cast<char*>(c->b_xty.p) + c->b_xty_stride * s);
    dev[s].gd.yty0 = (R)c->b_yty[s];
    dev[s].n_obs = c->b_nobs[s];
  }
  CU_TRY(c->b_dev.reserve(dev.size() * sizeof(ci::BatchDev<R>)));
  CU_TRY(cudaMemcpyAsync(c->b_dev.p, dev.data(), dev.size() * sizeof(ci::BatchDev<R>),
                         cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

}  // namespace

std::string& cih_err() {
  static thread_local std::string err;
  return err;
}

extern "C" {

int ci_version(void) { return CI_B200_VERSION; }
const char* ci_last_error(void) { return cih_err().c_str(); }

int ci_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { fail(CI_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return CI_ERR_NO_DEVICE; }
  return n;
}

int ci_ctx_create(int device, ci_ctx** out) {
  if (!out) return fail(CI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(CI_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(CI_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CI_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  ci_ctx* c = new ci_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  for (DevBuf* b : c->bufs()) b->retired = &c->retired;
  cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (se != cudaSuccess) { delete c; return fail(CI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(se)); }
  if (const char* g = getenv("CI_B200_G")) c->force_G = atoi(g);
  if (const char* g = getenv("CI_B200_TEAM")) c->team_mode = atoi(g);
  if (const char* g = getenv("CI_B200_PREDICT_TEAM")) c->predict_team = atoi(g);
  if (const char* g = getenv("CI_B200_TSTREAM")) c->tstream_mode = atoi(g);
  if (const char* g = getenv("CI_B200_GIBBS_TEAM")) c->gibbs_team = atoi(g);
  if (const char* g = getenv("CI_B200_SEL_NT")) c->sel_nt = atoi(g);
  if (const char* g = getenv("CI_B200_SEL_SMEM")) c->sel_smem = atoi(g);
  if (const char* g = getenv("CI_B200_TSW")) c->tstream_W = atoi(g);
  *out = c;
  return CI_OK;
}

int ci_ctx_destroy(ci_ctx* c) {
  if (!c) return CI_OK;
  cudaSetDevice(c->device);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  for (DevBuf* b : c->bufs()) b->release();
  c->free_retired();
  c->ring.release();
  delete c;
  return CI_OK;
}

int64_t ci_launch_count(const ci_ctx* c) { return c ? c->launches : 0; }

int ci_set_data(ci_ctx* c, const ci_problem* pb, const void* y, const void* X, const void* Omega) {
  if (!c || !pb || !y) return fail(CI_ERR_INVALID, "null argument");
  if (pb->T < 1) return fail(CI_ERR_INVALID, "T must be >= 1 (got %d)", pb->T);
  if (pb->p < 0) return fail(CI_ERR_INVALID, "p must be >= 0 (got %d)", pb->p);
  if (pb->p > 0 && (!X || !Omega)) return fail(CI_ERR_INVALID, "X and Omega are required when p > 0");
  if (pb->dtype != CI_F32 && pb->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (pb->model != CI_MODEL_LOCAL_LEVEL && pb->model != CI_MODEL_LOCAL_LINEAR_TREND)
    return fail(CI_ERR_INVALID, "unknown model %d", pb->model);
  if (pb->model == CI_MODEL_LOCAL_LINEAR_TREND && !(pb->P0_slope > 0))
    return fail(CI_ERR_INVALID, "P0_slope must be positive");
  const int d = pb->model == CI_MODEL_LOCAL_LINEAR_TREND ? 2 : 1;
  if (pb->p + 1 + d > ci::MAX_DIM)
    return fail(CI_ERR_UNSUPPORTED, "p=%d exceeds the supported maximum %d", pb->p, ci::MAX_DIM - 1 - d);
  if (!(pb->P0 > 0)) return fail(CI_ERR_INVALID, "P0 must be positive");
  CU_TRY(cudaSetDevice(c->device));
  c->free_retired();
  c->has_data = false;
  c->seas = ci::SeasDev{};
  c->prob = *pb;
  c->esz = pb->dtype == CI_F64 ? 8 : 4;
  c->NB = (pb->T + ci::TB - 1) / ci::TB;
  c->ld = ci::tile_ld(pb->p);
  c->dim = pb->p + 1 + d;
  const size_t te = (size_t)ci::tile_elems(pb->p);
  const size_t tile_bytes = (size_t)c->NB * te * c->esz;
  CU_TRY(c->tiles.reserve(tile_bytes));
  if (pb->dtype == CI_F64) {
    std::vector<double> h; build_tiles<double>(pb, y, X, c->NB, c->ld, h);
    CU_TRY(cudaMemcpyAsync(c->tiles.p, h.data(), tile_bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  } else {
    std::vector<float> h; build_tiles<float>(pb, y, X, c->NB, c->ld, h);
    CU_TRY(cudaMemcpyAsync(c->tiles.p, h.data(), tile_bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  const size_t ob = (size_t)pb->p * pb->p * c->esz;
  CU_TRY(c->omega.reserve(ob + 16));          // bulk copies move whole 16-byte units
  if (ob) {
    CU_TRY(cudaMemcpyAsync(c->omega.p, Omega, ob, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  {  // sufficient statistics of the regression step over OBSERVED rows (float64 on the host)
    const int T = pb->T, p = pb->p;
    std::vector<double> gram((size_t)p * p, 0.0), xty((size_t)(p > 0 ? p : 1), 0.0);
    double yty = 0.0; int nobs = 0;
    for (int t = 0; t < T; ++t) {
      const double yt = pb->dtype == CI_F64 ? static_cast<const double*>(y)[t]
                                            : (double)static_cast<const float*>(y)[t];
      if (!(yt == yt)) continue;
      ++nobs; yty += yt * yt;
      for (int i = 0; i < p; ++i) {
        const double xi = pb->dtype == CI_F64 ? static_cast<const double*>(X)[(size_t)t * p + i]
                                              : (double)static_cast<const float*>(X)[(size_t)t * p + i];
        xty[i] += xi * yt;
        for (int j = 0; j <= i; ++j) {
          const double xj = pb->dtype == CI_F64 ? static_cast<const double*>(X)[(size_t)t * p + j]
                                                : (double)static_cast<const float*>(X)[(size_t)t * p + j];
          gram[(size_t)i * p + j] += xi * xj;
        }
      }
    }
    for (int i = 0; i < p; ++i)
      for (int j = i + 1; j < p; ++j) gram[(size_t)i * p + j] = gram[(size_t)j * p + i];
    c->yty0 = yty; c->n_obs = nobs;
    CU_TRY(c->gram.reserve((size_t)p * p * c->esz + 16));
    CU_TRY(c->xty0.reserve((size_t)p * c->esz + 16));
    if (p > 0) {
      if (pb->dtype == CI_F64) {
        CU_TRY(cudaMemcpyAsync(c->gram.p, gram.data(), (size_t)p * p * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->xty0.p, xty.data(), (size_t)p * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
      } else {
        std::vector<float> gf(gram.begin(), gram.end()), xf(xty.begin(), xty.end());
        CU_TRY(cudaMemcpyAsync(c->gram.p, gf.data(), (size_t)p * p * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->xty0.p, xf.data(), (size_t)p * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
      }
    }
  }
  c->v_tiles = c->tiles.p; c->v_omega = c->omega.p; c->v_gram = c->gram.p; c->v_xty = c->xty0.p;
  c->batch_n = 0;
  // validate that the pipeline fits before accepting the problem
  ci::SmemCfg cfg;
  int rc = plan_smem(c, 1, 0, &cfg);
  if (rc) return rc;
  c->has_data = true;
  return CI_OK;
}

// ---- batches of independent series -------------------------------------
int ci_batch_select(ci_ctx* c, int s) {
  if (!c) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  if (s < 0 || s >= c->batch_n) return fail(CI_ERR_INVALID, "series %d out of range [0,%d)", s, c->batch_n);
  c->prob = c->b_prob[s];
  c->v_tiles = static_cast<char*>(c->b_tiles.p) + c->b_tile_stride * s;
  c->v_omega = static_cast<char*>(c->b_omega.p) + c->b_omega_stride * s;
  c->v_gram = static_cast<char*>(c->b_gram.p) + c->b_gram_stride * s;
  c->v_xty = static_cast<char*>(c->b_xty.p) + c->b_xty_stride * s;
  c->yty0 = c->b_yty[s]; c->n_obs = c->b_nobs[s];
  // (a season calendar set with ci_set_seasonal after ci_set_data_batch belongs to the whole
  // panel -- same T for every series -- and survives the selection)
  c->has_data = true;
  return CI_OK;
}

int ci_set_data_batch(ci_ctx* c, const ci_problem* probs, int N, const void* y, const void* X,
                      const void* Omega) {
  if (!c || !probs || !y) return fail(CI_ERR_INVALID, "null argument");
  if (N < 1) return fail(CI_ERR_INVALID, "n_series must be >= 1");
  const ci_problem& p0 = probs[0];
  if (p0.T < 1 || p0.p < 0) return fail(CI_ERR_INVALID, "bad T / p");
  if (p0.p > 0 && (!X || !Omega)) return fail(CI_ERR_INVALID, "X and Omega are required when p > 0");
  if (p0.dtype != CI_F32 && p0.dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (p0.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "batches are local-level only (as the reference's model)");
  if (p0.p + 2 > ci::MAX_DIM) return fail(CI_ERR_UNSUPPORTED, "p=%d exceeds the supported maximum", p0.p);
  for (int s = 0; s < N; ++s) {
    if (probs[s].T != p0.T || probs[s].p != p0.p || probs[s].dtype != p0.dtype || probs[s].model != p0.model)
      return fail(CI_ERR_INVALID, "series %d differs in T / p / dtype / model: a batch shares its shape", s);
    if (!(probs[s].P0 > 0)) return fail(CI_ERR_INVALID, "P0 must be positive (series %d)", s);
  }
  CU_TRY(cudaSetDevice(c->device));
  c->free_retired();
  c->has_data = false;
  c->seas = ci::SeasDev{};
  c->prob = p0;
  c->esz = p0.dtype == CI_F64 ? 8 : 4;
  c->NB = (p0.T + ci::TB - 1) / ci::TB;
  c->ld = ci::tile_ld(p0.p);
  c->dim = p0.p + 2;
  int rc = p0.dtype == CI_F64 ? upload_batch<double>(c, probs, N, y, X, Omega)
                              : upload_batch<float>(c, probs, N, y, X, Omega);
  if (rc) return rc;
  c->batch_n = N;
  rc = ci_batch_select(c, 0);
  if (rc) return rc;
  ci::SmemCfg cfg;
  rc = plan_smem(c, 1, 0, &cfg);
  if (rc) { c->has_data = false; c->batch_n = 0; return rc; }
  return CI_OK;
}

int ci_set_seasonal_batch(ci_ctx* c, const ci_seasonal* sp, const double* init_sd,
                          const double* drift_scale, const double* drift_ub) {
  if (!c || !sp || !init_sd || !drift_scale || !drift_ub) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  int rc = ci_set_seasonal(c, sp);
  if (rc) return rc;
  const int N = c->batch_n;
  std::vector<double> h((size_t)3 * N);
  for (int s = 0; s < N; ++s) {
    if (!(init_sd[s] > 0) || !(drift_scale[s] > 0) || !(drift_ub[s] > 0)) {
      c->seas = ci::SeasDev{};
      return fail(CI_ERR_INVALID, "init_sd, drift_scale, drift_ub must be positive (series %d)", s);
    }
    h[3 * s] = init_sd[s] * init_sd[s]; h[3 * s + 1] = drift_scale[s];
    h[3 * s + 2] = ub_variance(drift_ub[s], c->b_prob[s].ub_on_scale);
  }
  CU_TRY(c->s_series.reserve(h.size() * sizeof(double)));
  CU_TRY(cudaMemcpyAsync(c->s_series.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->seas.per_series = static_cast<const double*>(c->s_series.p);
  return CI_OK;
}

int ci_set_seasonal(ci_ctx* c, const ci_seasonal* sp) {
  if (!c) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  c->seas = ci::SeasDev{};
  if (!sp || sp->n_components == 0) return CI_OK;
  const int K = sp->n_components, T = c->prob.T;
  if (K < 0 || K > CI_MAX_SEASONAL)
    return fail(CI_ERR_UNSUPPORTED, "at most %d seasonal components (got %d)", CI_MAX_SEASONAL, K);
  if (c->prob.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "seasonal components need the local level model");
  if (!sp->active || !sp->ends) return fail(CI_ERR_INVALID, "null schedule");
  if (!(sp->init_sd > 0) || !(sp->drift_conc > 0) || !(sp->drift_scale > 0) || !(sp->drift_ub > 0))
    return fail(CI_ERR_INVALID, "init_sd, drift_conc, drift_scale, drift_ub must be positive");
  ci::SeasDev sz{};
  sz.K = K;
  int d = 1;
  for (int k = 0; k < K; ++k) {
    if (sp->num_seasons[k] < 2) return fail(CI_ERR_INVALID, "num_seasons must be >= 2");
    sz.n[k] = sp->num_seasons[k]; sz.off[k] = d; d += sz.n[k];
  }
  if (d > ci::SEAS_MAXD)
    return fail(CI_ERR_UNSUPPORTED, "1 + sum(num_seasons) = %d exceeds the supported state "
                "dimension %d", d, ci::SEAS_MAXD);
  sz.d = d;
  std::vector<uint8_t> sched((size_t)T * (K + 1));
  for (int t = 0; t < T; ++t) {
    uint8_t em = 0;
    for (int k = 0; k < K; ++k) {
      const uint8_t a = sp->active[(size_t)k * T + t];
      if (a >= sz.n[k]) return fail(CI_ERR_INVALID, "active[%d][%d] = %d out of range", k, t, (int)a);
      sched[(size_t)t * (K + 1) + k] = a;
      if (sp->ends[(size_t)k * T + t]) { em |= (uint8_t)(1u << k); if (t < T - 1) sz.n_ends[k]++; }
    }
    sched[(size_t)t * (K + 1) + K] = em;
  }
  sz.init_var = sp->init_sd * sp->init_sd;
  sz.drift_conc = sp->drift_conc; sz.drift_scale = sp->drift_scale; sz.drift_ub = ub_variance(sp->drift_ub, c->prob.ub_on_scale);
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(c->s_sched.reserve(sched.size()));
  CU_TRY(cudaMemcpyAsync(c->s_sched.p, sched.data(), sched.size(), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  sz.sched = static_cast<const uint8_t*>(c->s_sched.p);
  c->seas = sz;
  return CI_OK;
}

}  // extern "C"
