// shared by the translation units of the reference-named C API (capi_pgm.cpp, capi_pgm_meta.cpp): the handle and the
// Lippincott wrapper (power_grid_model_c/src/handle.hpp:20-89)
#pragma once

#include "../../include/pgm_b200_capi.h"
#include "pgm_meta.hpp"

#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

struct PGM_Handle {
    PGM_Idx err_code{PGM_no_error};
    std::string err_msg;
    std::vector<PGM_Idx> failed_scenarios;
    std::vector<std::string> batch_errs;
    mutable std::vector<char const*> batch_errs_c_str;
};

namespace pgmb::capi {

inline void clear(PGM_Handle* handle) {
    if (handle != nullptr) *handle = PGM_Handle{};
}

// call_with_catch: clear the handle, run, translate any exception into PGM_regular_error + message
template <class F> auto call(PGM_Handle* handle, F&& f) noexcept -> decltype(f()) {
    using R = decltype(f());
    try {
        clear(handle);
        return f();
    } catch (std::exception const& e) {
        if (handle != nullptr) {
            handle->err_code = PGM_regular_error;
            handle->err_msg = e.what();
        }
    } catch (...) {
        if (handle != nullptr) {
            handle->err_code = PGM_regular_error;
            handle->err_msg = "Unknown error!\n";
        }
    }
    if constexpr (!std::is_void_v<R>) return R{};
}

template <class T> T& deref(T* p) {
    if (p == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
    return *p;
}

} // namespace pgmb::capi
