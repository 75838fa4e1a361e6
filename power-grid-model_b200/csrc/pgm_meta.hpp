// Meta-data tables of the reference's datasets (auxiliary/meta_data.hpp:95-200): dataset -> component -> attribute with
// name, ctype and offset; sizes and alignments of the component structs.  The table itself (meta_table.inc) is generated from
// the reference's definition files by tools/gen_meta_table.py.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string_view>

struct PGM_MetaAttribute {
    char const* name;
    int32_t ctype; // PGM_CType: 0 int32, 1 int8, 2 double, 3 double[3]
    size_t offset;
    size_t component_size;
    size_t size() const { return ctype == 0 ? 4 : ctype == 1 ? 1 : ctype == 2 ? 8 : 24; }
};
struct PGM_MetaComponent {
    char const* name;
    size_t size;
    size_t alignment;
    int64_t n_attributes;
    PGM_MetaAttribute const* attributes;
    PGM_MetaAttribute const* find(std::string_view attribute) const; // nullptr when absent
    void set_nan(void* buffer, int64_t begin, int64_t count) const;
};
struct PGM_MetaDataset {
    char const* name;
    int64_t n_components;
    PGM_MetaComponent const* components;
    PGM_MetaComponent const* find(std::string_view component) const;
};

namespace pgmb::meta {
int64_t n_datasets();
PGM_MetaDataset const* dataset(int64_t idx);
PGM_MetaDataset const* find_dataset(std::string_view name); // nullptr when absent
} // namespace pgmb::meta
