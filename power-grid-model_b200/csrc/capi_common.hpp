// shared helpers of the C-ABI translation units
#pragma once

#include "../../include/pgm_b200.h"
#include "engine.hpp"

#include <string>

namespace pgmb {

extern thread_local std::string g_last_error;
MathTopology topology_from_view(pgmb_math_topology const& t);

template <class F> int guarded(F&& f) noexcept {
    try {
        g_last_error.clear();
        f();
        return PGMB_OK;
    } catch (CudaError const& e) {
        g_last_error = e.what();
        return PGMB_ERR_CUDA;
    } catch (std::invalid_argument const& e) {
        g_last_error = e.what();
        return PGMB_ERR_INVALID;
    } catch (InvalidArgument const& e) {
        g_last_error = e.what();
        return PGMB_ERR_INVALID;
    } catch (std::exception const& e) {
        g_last_error = e.what();
        return PGMB_ERR_INTERNAL;
    } catch (...) {
        g_last_error = "unknown error";
        return PGMB_ERR_INTERNAL;
    }
}

} // namespace pgmb
