#include "common.h"

#include <string.h>

#include "../../include/osudit.h"

namespace osudit {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}

static int g_sm_limit = 0;  // 0 = none; set by osudit_set_sm_limit

static int device_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

int num_sms() {
  const int n = device_sms();
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

}  // namespace osudit

extern "C" const char* osudit_last_error(void) { return osudit::g_err; }
extern "C" int osudit_version(void) { return OSUDIT_VERSION; }
extern "C" int osudit_set_sm_limit(int n) {
  if (n == -1) return osudit::g_sm_limit;  // query
  if (n < 0) return osudit::set_error(-2, "set_sm_limit: negative limit");
  const int prev = osudit::g_sm_limit;
  osudit::g_sm_limit = n & ~1;  // CTA pairs: keep it even
  return prev;
}
