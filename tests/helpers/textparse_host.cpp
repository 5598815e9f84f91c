// Host build of jax_powspec_b200/csrc/textparse.cuh (the field / line converters the device reader
// runs) so that the CPU test-suite can fuzz them against strtod / np.loadtxt without a GPU.
// Test infrastructure only: nothing in the package loads this.
#include "../../jax_powspec_b200/csrc/textparse.cuh"

extern "C" {

int jps_host_parse_field(const char* s, int len, int comment, float* out, int* consumed) {
  const char* next = s;
  const int st = jps::text::parse_field(s, s + len, (char)comment, *out, next);
  *consumed = (int)(next - s);
  return st;
}

// returns 1 if the line holds a row, 0 if it is blank / comment only; *status = worst field status
int jps_host_parse_line(const char* s, int len, int comment, const int* cols, int ncols, float* vals, int* status) {
  return jps::text::parse_line<8>(s, s + len, (char)comment, cols, ncols, vals, *status);
}

// batch: n NUL-free fields packed back to back, offsets[n+1]
void jps_host_parse_fields(const char* buf, const long long* offsets, int n, float* out, int* status) {
  for (int i = 0; i < n; ++i) {
    const char* next;
    status[i] = jps::text::parse_field(buf + offsets[i], buf + offsets[i + 1], '#', out[i], next);
    if (status[i] == jps::text::FIELD_OK && next != buf + offsets[i + 1]) status[i] = 100;   // must consume all
  }
}
}
