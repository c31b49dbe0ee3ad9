// shim (test infrastructure): LOG(severity) << ... swallowed
#pragma once
#include <gflags/gflags.h>
#include <iostream>
namespace lvo_shim { struct NullLog { template <class T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define LOG(sev) lvo_shim::NullLog()
#define CHECK(x) if (!(x)) abort(); else lvo_shim::NullLog()
