// Version / error plumbing of the C ABI (include/idb200.h).
#include "common.cuh"

namespace idb200 {

char* last_error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace idb200

extern "C" int idb200_version(void) { return 100; }  // 0.1.0
extern "C" const char* idb200_last_error(void) { return idb200::last_error_buffer(); }
