// Internal helpers shared by the translation units that implement the C ABI.
#pragma once
#include <string>

namespace glr {
int api_fail(int code, const std::string& msg);  // records the thread's last error, returns `code`
int api_check_device(int device);                // validates the ordinal (sm_100) and makes it current
}  // namespace glr
