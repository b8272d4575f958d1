// Thread-local error string behind eucl_last_error().
#pragma once
#include <string>

namespace eucl {
void set_last_error(const std::string& message);
int fail(int status, const std::string& message); // records the message, returns status
} // namespace eucl
