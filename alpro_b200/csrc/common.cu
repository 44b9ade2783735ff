// Library-wide host utilities: last-error string, device properties.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.h"

namespace alpro {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  });
  return n;
}

bool pdl_enabled() {   // read per launch (~50 ns) so that a test or the bench can switch it inside one process
  const char* e = getenv("ALPRO_PDL");
  return !(e && e[0] == '0');
}

}  // namespace alpro

extern "C" const char* alpro_last_error(void) { return alpro::g_err; }
extern "C" int alpro_version(void) { return 1; }
extern "C" int alpro_num_sms(void) { return alpro::num_sms(); }
