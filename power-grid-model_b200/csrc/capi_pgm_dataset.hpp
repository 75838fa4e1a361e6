// Dataset views of the reference-named C API (power_grid_model_c/src/dataset.cpp, auxiliary/dataset.hpp): shared by
// capi_pgm.cpp (const / mutable datasets, PGM_calculate) and capi_pgm_serialization.cpp (writable datasets of the deserializer).
#pragma once

#include "capi_pgm_common.hpp"
#include "engine.hpp"

#include <memory>
#include <string>
#include <vector>

namespace pgmb::capi {

struct DatasetError : std::runtime_error {
    explicit DatasetError(std::string const& msg) : std::runtime_error("Dataset error: " + msg) {}
};
struct AttributeBuffer {
    PGM_MetaAttribute const* attribute;
    void* data;
};

struct DatasetBuffer {
    std::string component;
    PGM_Idx elements_per_scenario;
    PGM_Idx total_elements;
    PGM_Idx const* indptr;
    void* data; // row buffer; nullptr = columnar component, its attribute buffers are in `attributes`
    PGM_MetaComponent const* meta;
    std::vector<AttributeBuffer> attributes;
    // attribute indications (auxiliary/dataset.hpp: ComponentInfo::has_attribute_indications / attribute_indications): set by
    // the deserializer when every row of the component is a compact list, so that a client can pick columnar buffers
    bool has_indications{false};
    std::vector<PGM_MetaAttribute const*> indications;
    bool columnar() const { return data == nullptr; }
};

struct Dataset {
    std::string name;
    bool is_batch;
    PGM_Idx batch_size;
    PGM_MetaDataset const* meta;
    std::vector<DatasetBuffer> buffers;
    Dataset const* next{nullptr}; // next cartesian-product dimension

    Dataset(char const* dataset, PGM_Idx batch, PGM_Idx size) : is_batch{batch != 0}, batch_size{size} {
        if (dataset == nullptr) throw InvalidArgument("Received null pointer where a valid pointer was expected.\n");
        name = dataset;
        meta = meta::find_dataset(name);
        if (meta == nullptr) throw std::out_of_range("Cannot find dataset with name: " + name + "!\n");
        if (batch_size < 0) throw DatasetError("Batch size cannot be negative!\n");
        if (!is_batch && batch_size != 1) throw DatasetError("For non-batch dataset, batch size should be one!\n");
    }

    DatasetBuffer* find(std::string const& component) {
        for (auto& b : buffers) {
            if (b.component == component) return &b;
        }
        return nullptr;
    }
    DatasetBuffer const& at(PGM_Idx idx) const {
        if (idx < 0 || idx >= static_cast<PGM_Idx>(buffers.size())) throw std::out_of_range("Index out of range!\n");
        return buffers[static_cast<size_t>(idx)];
    }

    // auxiliary/dataset.hpp:587-625
    void add_buffer(char const* component, PGM_Idx elements_per_scenario, PGM_Idx total_elements, PGM_Idx const* indptr,
                    void* data, bool check_indptr) {
        if (component == nullptr) throw InvalidArgument("Received null pointer where a valid pointer was expected.\n");
        PGM_MetaComponent const* mc = meta->find(component);
        if (mc == nullptr) throw std::out_of_range("Cannot find component with name: " + std::string(component) + "!\n");
        if (find(component) != nullptr) throw DatasetError("Cannot have duplicated components!\n");
        if (elements_per_scenario >= 0 && elements_per_scenario * batch_size != total_elements) {
            throw DatasetError("For a uniform buffer, total_elements should be equal to elements_per_scenario * batch_size!\n");
        }
        if (elements_per_scenario < 0) {
            if (indptr == nullptr) throw DatasetError("For a non-uniform buffer, indptr should be supplied!\n");
            if (check_indptr) {
                if (indptr[0] != 0 || indptr[batch_size] != total_elements) {
                    throw DatasetError("For a non-uniform buffer, indptr should begin with 0 and end with total_elements!\n");
                }
                for (PGM_Idx s = 0; s != batch_size; ++s) {
                    if (indptr[s] > indptr[s + 1]) throw DatasetError("For a non-uniform buffer, indptr should be non-decreasing!\n");
                }
            }
        } else if (indptr != nullptr) {
            throw DatasetError("For a uniform buffer, indptr should be nullptr!\n");
        }
        buffers.push_back({component, elements_per_scenario, total_elements, indptr, data, mc, {}});
    }

    // auxiliary/dataset.hpp:563-574
    void set_next(Dataset const* next_dataset) {
        for (Dataset const* d = next_dataset; d != nullptr; d = d->next) {
            if (d == this) throw DatasetError("Cannot create cyclic cartesian product dimension linked list!\n");
        }
        next = next_dataset;
    }

    // rows [begin, end) of a buffer as a buffer of its own (row pointer or every attribute pointer moved)
    static DatasetBuffer sub_buffer(DatasetBuffer const& b, PGM_Idx begin, PGM_Idx end, PGM_Idx elements_per_scenario) {
        DatasetBuffer r{b.component, elements_per_scenario, end - begin, nullptr, nullptr, b.meta, {}};
        if (!b.columnar()) r.data = static_cast<char*>(b.data) + static_cast<size_t>(begin) * b.meta->size;
        for (auto const& a : b.attributes) {
            r.attributes.push_back({a.attribute, static_cast<char*>(a.data) + static_cast<size_t>(begin) * a.attribute->size()});
        }
        return r;
    }
    // get_individual_scenario (auxiliary/dataset.hpp:463-475): scenario i as a single (non-batch) dataset
    Dataset individual_scenario(PGM_Idx i) const {
        Dataset single{name.c_str(), 0, 1};
        for (auto const& b : buffers) {
            PGM_Idx const begin = b.elements_per_scenario < 0 ? b.indptr[i] : i * b.elements_per_scenario;
            PGM_Idx const end = b.elements_per_scenario < 0 ? b.indptr[i + 1] : (i + 1) * b.elements_per_scenario;
            single.buffers.push_back(sub_buffer(b, begin, end, end - begin));
        }
        return single;
    }
    // get_slice_scenario (auxiliary/dataset.hpp:476-499): scenarios [begin, end) of a batch dataset with uniform buffers
    Dataset slice_scenarios(PGM_Idx begin, PGM_Idx end) const {
        Dataset slice{name.c_str(), 1, end - begin};
        for (auto const& b : buffers) {
            if (b.elements_per_scenario < 0) throw DatasetError("Cannot export a single dataset with specified scenario\n");
            slice.buffers.push_back(sub_buffer(b, begin * b.elements_per_scenario, end * b.elements_per_scenario, b.elements_per_scenario));
        }
        return slice;
    }

    // auxiliary/dataset.hpp:627-646: a component added with a null row pointer is columnar and takes one buffer per attribute
    void add_attribute_buffer(char const* component, char const* attribute, void* data) {
        if (component == nullptr || attribute == nullptr) throw InvalidArgument("Received null pointer where a valid pointer was expected.\n");
        DatasetBuffer* b = find(component);
        if (b == nullptr) throw DatasetError("Cannot find component '" + std::string(component) + "'!\n");
        if (!b->columnar()) throw DatasetError("Cannot add attribute buffers to row-based dataset!\n");
        PGM_MetaAttribute const* ma = b->meta->find(attribute);
        if (ma == nullptr) throw std::out_of_range("Cannot find attribute with name: " + std::string(attribute) + "!\n");
        for (auto const& a : b->attributes) {
            if (a.attribute == ma) throw DatasetError("Cannot have duplicated attribute buffers!\n");
        }
        if (data == nullptr && b->total_elements > 0) {
            throw DatasetError("Attribute buffer data pointer cannot be null for non-empty component!\n");
        }
        b->attributes.push_back({ma, data});
    }
};


} // namespace pgmb::capi

// the opaque dataset types of dataset.h: three views of the same structure
struct PGM_ConstDataset : pgmb::capi::Dataset {
    using Dataset::Dataset;
};
struct PGM_MutableDataset : pgmb::capi::Dataset {
    using Dataset::Dataset;
};
struct PGM_WritableDataset : pgmb::capi::Dataset {
    using Dataset::Dataset;
};
