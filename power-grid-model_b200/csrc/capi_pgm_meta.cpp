// PGM_meta_* (power_grid_model_c/src/meta_data.cpp:28-160) and PGM_*buffer* (src/buffer.cpp:33-108) of the reference's C API:
// what a client needs to size, fill and read the row buffers it hands to PGM_create_model / PGM_calculate without hard-coding
// struct layouts (the reference's Python wrapper builds its numpy dtypes from exactly these calls, _core/power_grid_meta.py).
#include "../../include/pgm_b200_capi.h"
#include "capi_pgm_common.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <unordered_set>

#include "../../include/pgm_b200.h"

using MetaAttribute = PGM_MetaAttribute;
using MetaComponent = PGM_MetaComponent;
using MetaDataset = PGM_MetaDataset;
namespace {
#include "meta_table.inc"

// the PGM_def_* pointer constants of dataset_definitions.h (declared with C linkage and default visibility in the header)
#include "../../include/pgm_b200_dataset_definitions.h"
extern "C" {
#include "meta_defs.inc"
}
} // namespace

PGM_MetaAttribute const* PGM_MetaComponent::find(std::string_view attribute) const {
    for (int64_t i = 0; i != n_attributes; ++i) {
        if (attribute == attributes[i].name) return attributes + i;
    }
    return nullptr;
}
PGM_MetaComponent const* PGM_MetaDataset::find(std::string_view component) const {
    for (int64_t i = 0; i != n_components; ++i) {
        if (component == components[i].name) return components + i;
    }
    return nullptr;
}
void PGM_MetaComponent::set_nan(void* buffer, int64_t begin, int64_t count) const {
    // null values of the reference: na_IntID = INT32_MIN, na_IntS = INT8_MIN, NaN (common/common.hpp:87-90)
    double const nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t i = begin; i != begin + count; ++i) {
        char* row = static_cast<char*>(buffer) + static_cast<size_t>(i) * size;
        for (int64_t a = 0; a != n_attributes; ++a) {
            char* p = row + attributes[a].offset;
            switch (attributes[a].ctype) {
            case 0: { int32_t const v = std::numeric_limits<int32_t>::min(); std::memcpy(p, &v, 4); break; }
            case 1: { int8_t const v = std::numeric_limits<int8_t>::min(); std::memcpy(p, &v, 1); break; }
            case 2: std::memcpy(p, &nan, 8); break;
            default: for (int k = 0; k != 3; ++k) std::memcpy(p + 8 * k, &nan, 8);
            }
        }
    }
}

namespace pgmb::meta {
int64_t n_datasets() { return kNDatasets; }
PGM_MetaDataset const* dataset(int64_t idx) { return idx >= 0 && idx < kNDatasets ? kDatasets + idx : nullptr; }
PGM_MetaDataset const* find_dataset(std::string_view name) {
    for (auto const& d : kDatasets) {
        if (name == d.name) return &d;
    }
    return nullptr;
}
} // namespace pgmb::meta

using namespace pgmb;
using namespace pgmb::capi;

namespace {
std::mutex g_locked_mutex;
std::unordered_set<void*> g_locked_buffers; // page-locked buffers handed out by PGM_create_buffer
} // namespace

namespace {
MetaDataset const& dataset_by_name(char const* name) {
    if (name == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
    auto const* d = meta::find_dataset(name);
    if (d == nullptr) throw std::out_of_range("Cannot find dataset with name: " + std::string(name) + "!\n");
    return *d;
}
MetaComponent const& component_by_name(char const* dataset, char const* component) {
    MetaDataset const& d = dataset_by_name(dataset);
    if (component == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
    auto const* c = d.find(component);
    if (c == nullptr) throw std::out_of_range("Cannot find component with name: " + std::string(component) + "!\n");
    return *c;
}
} // namespace

extern "C" {

PGM_Idx PGM_meta_n_datasets(PGM_Handle* handle) {
    return call(handle, [] { return static_cast<PGM_Idx>(meta::n_datasets()); });
}
PGM_MetaDataset const* PGM_meta_get_dataset_by_idx(PGM_Handle* handle, PGM_Idx idx) {
    return call(handle, [&] {
        auto const* d = meta::dataset(idx);
        if (d == nullptr) throw std::out_of_range("Index out of range!\n");
        return d;
    });
}
PGM_MetaDataset const* PGM_meta_get_dataset_by_name(PGM_Handle* handle, char const* dataset) {
    return call(handle, [&] { return &dataset_by_name(dataset); });
}
char const* PGM_meta_dataset_name(PGM_Handle* handle, PGM_MetaDataset const* dataset) {
    return call(handle, [&] { return deref(dataset).name; });
}
PGM_Idx PGM_meta_n_components(PGM_Handle* handle, PGM_MetaDataset const* dataset) {
    return call(handle, [&] { return static_cast<PGM_Idx>(deref(dataset).n_components); });
}
PGM_MetaComponent const* PGM_meta_get_component_by_idx(PGM_Handle* handle, PGM_MetaDataset const* dataset, PGM_Idx idx) {
    return call(handle, [&] {
        MetaDataset const& d = deref(dataset);
        if (idx < 0 || idx >= d.n_components) throw std::out_of_range("Index out of range!\n");
        return d.components + idx;
    });
}
PGM_MetaComponent const* PGM_meta_get_component_by_name(PGM_Handle* handle, char const* dataset, char const* component) {
    return call(handle, [&] { return &component_by_name(dataset, component); });
}
char const* PGM_meta_component_name(PGM_Handle* handle, PGM_MetaComponent const* component) {
    return call(handle, [&] { return deref(component).name; });
}
size_t PGM_meta_component_size(PGM_Handle* handle, PGM_MetaComponent const* component) {
    return call(handle, [&] { return deref(component).size; });
}
size_t PGM_meta_component_alignment(PGM_Handle* handle, PGM_MetaComponent const* component) {
    return call(handle, [&] { return deref(component).alignment; });
}
PGM_Idx PGM_meta_n_attributes(PGM_Handle* handle, PGM_MetaComponent const* component) {
    return call(handle, [&] { return static_cast<PGM_Idx>(deref(component).n_attributes); });
}
PGM_MetaAttribute const* PGM_meta_get_attribute_by_idx(PGM_Handle* handle, PGM_MetaComponent const* component, PGM_Idx idx) {
    return call(handle, [&] {
        MetaComponent const& c = deref(component);
        if (idx < 0 || idx >= c.n_attributes) throw std::out_of_range("Index out of range!\n");
        return c.attributes + idx;
    });
}
PGM_MetaAttribute const* PGM_meta_get_attribute_by_name(PGM_Handle* handle, char const* dataset, char const* component,
                                                        char const* attribute) {
    return call(handle, [&] {
        MetaComponent const& c = component_by_name(dataset, component);
        if (attribute == nullptr) throw std::invalid_argument("Received null pointer where a valid pointer was expected.\n");
        auto const* a = c.find(attribute);
        if (a == nullptr) throw std::out_of_range("Cannot find attribute with name: " + std::string(attribute) + "!\n");
        return a;
    });
}
char const* PGM_meta_attribute_name(PGM_Handle* handle, PGM_MetaAttribute const* attribute) {
    return call(handle, [&] { return deref(attribute).name; });
}
PGM_Idx PGM_meta_attribute_ctype(PGM_Handle* handle, PGM_MetaAttribute const* attribute) {
    return call(handle, [&] { return static_cast<PGM_Idx>(deref(attribute).ctype); });
}
size_t PGM_meta_attribute_offset(PGM_Handle* handle, PGM_MetaAttribute const* attribute) {
    return call(handle, [&] { return deref(attribute).offset; });
}
int PGM_is_little_endian(PGM_Handle* handle) {
    return call(handle, [] {
        uint32_t const one = 1;
        unsigned char first;
        std::memcpy(&first, &one, 1);
        return static_cast<int>(first);
    });
}

// ---- buffers -------------------------------------------------------------------------------------------------------
// Buffers of at least 4 KB are page-locked when a CUDA device is present (cudaHostAlloc): a client that allocates its datasets
// through the API -- the reference's C++ wrapper and benchmark do (power_grid_model_cpp/buffer.hpp, fictional_grid_generator.hpp:
// 192-204) -- then gets the direct, chunk-overlapped transfers of the device pipeline without knowing about CUDA.  Elsewhere
// (no device, small buffers, allocation refused) it is the reference's aligned_alloc.
void* PGM_create_buffer(PGM_Handle* handle, PGM_MetaComponent const* component, PGM_Idx size) {
    return call(handle, [&]() -> void* {
        MetaComponent const& c = deref(component);
        size_t const alignment = std::max(c.alignment, sizeof(void*));
        size_t const bytes = c.size * static_cast<size_t>(std::max<PGM_Idx>(size, 0));
        if (bytes >= (size_t{4} << 10) && pgmb_device_count() > 0) {
            void* p = nullptr;
            if (pgmb_host_alloc(bytes, &p) == PGMB_OK && p != nullptr) {
                std::lock_guard<std::mutex> const lock(g_locked_mutex);
                g_locked_buffers.insert(p);
                return p;
            }
        }
        return std::aligned_alloc(alignment, (bytes + alignment - 1) / alignment * alignment);
    });
}
void PGM_destroy_buffer(void* ptr) {
    if (ptr == nullptr) return;
    {
        std::lock_guard<std::mutex> const lock(g_locked_mutex);
        auto const it = g_locked_buffers.find(ptr);
        if (it != g_locked_buffers.end()) {
            g_locked_buffers.erase(it);
            pgmb_host_free(ptr);
            return;
        }
    }
    std::free(ptr);
}
int PGM_b200_buffer_is_page_locked(void const* ptr) {
    std::lock_guard<std::mutex> const lock(g_locked_mutex);
    return g_locked_buffers.count(const_cast<void*>(ptr)) != 0 ? 1 : 0;
}
void PGM_buffer_set_nan(PGM_Handle* handle, PGM_MetaComponent const* component, void* ptr, PGM_Idx buffer_offset, PGM_Idx size) {
    call(handle, [&] { deref(component).set_nan(&deref(static_cast<char*>(ptr)), buffer_offset, size); });
}
void PGM_buffer_set_value(PGM_Handle* handle, PGM_MetaAttribute const* attribute, void* buffer_ptr, void const* src_ptr,
                          PGM_Idx buffer_offset, PGM_Idx size, PGM_Idx src_stride) {
    call(handle, [&] {
        MetaAttribute const& a = deref(attribute);
        if (size <= 0) return;
        PGM_Idx const stride = src_stride < 0 ? static_cast<PGM_Idx>(a.size()) : src_stride;
        char* rows = &deref(static_cast<char*>(buffer_ptr));
        char const* src = &deref(static_cast<char const*>(src_ptr));
        for (PGM_Idx i = buffer_offset; i != buffer_offset + size; ++i) { // the value pointer moves with the row index
            std::memcpy(rows + static_cast<size_t>(i) * a.component_size + a.offset, src + stride * i, a.size());
        }
    });
}
void PGM_buffer_get_value(PGM_Handle* handle, PGM_MetaAttribute const* attribute, void const* buffer_ptr, void* dest_ptr,
                          PGM_Idx buffer_offset, PGM_Idx size, PGM_Idx dest_stride) {
    call(handle, [&] {
        MetaAttribute const& a = deref(attribute);
        if (size <= 0) return;
        PGM_Idx const stride = dest_stride < 0 ? static_cast<PGM_Idx>(a.size()) : dest_stride;
        char const* rows = &deref(static_cast<char const*>(buffer_ptr));
        char* dest = &deref(static_cast<char*>(dest_ptr));
        for (PGM_Idx i = buffer_offset; i != buffer_offset + size; ++i) {
            std::memcpy(dest + stride * i, rows + static_cast<size_t>(i) * a.component_size + a.offset, a.size());
        }
    });
}

} // extern "C"
