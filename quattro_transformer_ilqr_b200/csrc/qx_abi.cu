// C-ABI plumbing: version, thread-local error text, layout queries (host-only functions).
#include <stdarg.h>

#include "qx_common.cuh"

namespace qx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return QX_ECUDA;
}

static bool dims(int model_id, int* n, int* m) {
  switch (model_id) {
    case QX_MODEL_CARTPOLE: *n = 4; *m = 1; return true;
    case QX_MODEL_QUADROTOR: *n = 12; *m = 4; return true;
  }
  return false;
}

}  // namespace qx

extern "C" int qx_abi_version(void) { return QX_ABI_VERSION; }

extern "C" const char* qx_last_error(void) { return qx::g_err; }

extern "C" int qx_layout(int model_id) {
  switch (model_id) {
    case QX_MODEL_CARTPOLE: return QX_LAYOUT_SOA;
    case QX_MODEL_QUADROTOR: return QX_LAYOUT_REC;
  }
  qx::set_error("unknown model_id %d", model_id);
  return QX_EINVAL;
}

extern "C" int qx_record_floats(int model_id) {
  int n, m;
  if (!qx::dims(model_id, &n, &m)) {
    qx::set_error("unknown model_id %d", model_id);
    return QX_EINVAL;
  }
  return 2 * n * n + 2 * n * m + n + m + m * m;
}

extern "C" size_t qx_expansion_bytes(int model_id, int B, int N) {
  const int r = qx_record_floats(model_id);
  if (r < 0 || B <= 0 || N <= 0) return 0;
  return (size_t)r * (size_t)B * (size_t)N * sizeof(float);
}

extern "C" size_t qx_terminal_bytes(int model_id, int B) {
  int n, m;
  if (!qx::dims(model_id, &n, &m) || B <= 0) return 0;
  return (size_t)(n + n * n) * (size_t)B * sizeof(float);
}

extern "C" size_t qx_gains_bytes(int model_id, int B, int N) {
  int n, m;
  if (!qx::dims(model_id, &n, &m) || B <= 0 || N <= 0) return 0;
  return (size_t)(m + m * n) * (size_t)B * (size_t)N * sizeof(float);
}
