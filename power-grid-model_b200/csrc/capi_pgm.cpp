// The reference's own C API names for the power-flow path (include/pgm_b200_capi.h): handle, options, row-based datasets,
// model create / update / copy / get_indexer / calculate.  A thin adapter: every call lands in the same pgmb::Model the
// pgmb_model_* seam drives, so PGM_calculate runs the CUDA pipeline and nothing else (no CPU fallback).
//   handle    power_grid_model_c/src/handle.cpp:29-62, handle.hpp:20-82 (error code / message / failed scenarios, cleared by
//             every call that takes the handle)
//   options   power_grid_model_c/src/options.cpp, options.hpp:16-27 (defaults)
//   datasets  power_grid_model_c/src/dataset.cpp:96-190 ; auxiliary/dataset.hpp:233-243, 587-625 (integrity checks, messages)
//   model     power_grid_model_c/src/model.cpp:48-75, 188-204 (single batch dimension), 349-359
#include "../../include/pgm_b200_capi.h"
#include "capi_common.hpp"
#include "capi_pgm_common.hpp"
#include "capi_pgm_dataset.hpp"
#include "model.hpp"

#include <cstdlib>
#include <cstring>

using namespace pgmb;
using namespace pgmb::capi;

struct PGM_Options {
    PGM_Idx calculation_type{PGM_power_flow};
    PGM_Idx calculation_method{PGM_default_method};
    PGM_Idx symmetric{1};
    double err_tol{1e-8};
    PGM_Idx max_iter{20};
    PGM_Idx threading{-1};
    PGM_Idx short_circuit_voltage_scaling{1};
    PGM_Idx tap_changing_strategy{0};
    PGM_Idx experimental_features{0};
};

namespace {

struct CalculationError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Row images of columnar components: the engine reads and writes packed rows, so a columnar input / update component is packed
// into a scratch row buffer (absent attributes = the reference's null values, like ColumnarAttributeRange's proxy does), and a
// columnar output component is computed into scratch rows whose requested attributes are copied out afterwards.
struct RowScratch {
    std::vector<std::unique_ptr<char[]>> store;
    struct Pending {
        DatasetBuffer const* buffer;
        char const* rows;
    };
    std::vector<Pending> outputs;

    void* rows_in(DatasetBuffer const& b) {
        if (!b.columnar()) return b.data;
        size_t const n = static_cast<size_t>(b.total_elements), size = b.meta->size;
        store.push_back(std::make_unique<char[]>(std::max<size_t>(n * size, 1)));
        char* rows = store.back().get();
        b.meta->set_nan(rows, 0, b.total_elements);
        for (auto const& a : b.attributes) {
            size_t const w = a.attribute->size();
            auto const* src = static_cast<char const*>(a.data);
            for (size_t i = 0; i != n; ++i) std::memcpy(rows + i * size + a.attribute->offset, src + i * w, w);
        }
        return rows;
    }
    void* rows_out(DatasetBuffer const& b) {
        if (!b.columnar()) return b.data;
        store.push_back(std::make_unique<char[]>(std::max<size_t>(static_cast<size_t>(b.total_elements) * b.meta->size, 1)));
        outputs.push_back({&b, store.back().get()});
        return store.back().get();
    }
    void scatter_outputs() const {
        for (auto const& p : outputs) {
            size_t const n = static_cast<size_t>(p.buffer->total_elements), size = p.buffer->meta->size;
            for (auto const& a : p.buffer->attributes) {
                size_t const w = a.attribute->size();
                auto* dst = static_cast<char*>(a.data);
                for (size_t i = 0; i != n; ++i) std::memcpy(dst + i * w, p.rows + i * size + a.attribute->offset, w);
            }
        }
    }
};

// power flow does not read these (sensors: state estimation; faults: short circuit); they may sit in the input dataset
bool ignored_by_power_flow(std::string const& c) {
    return c == "fault" || c.find("_sensor") != std::string::npos;
}

ComponentBuffer buffer_of(DatasetBuffer const& b, RowScratch& scratch) {
    if (b.elements_per_scenario < 0) return {-1, b.indptr, scratch.rows_in(b)};
    return {b.elements_per_scenario, nullptr, scratch.rows_in(b)};
}

int device_ordinal() {
    char const* env = std::getenv("PGMB_DEVICE");
    return env != nullptr ? std::atoi(env) : 0;
}

UpdateData update_of(Dataset const& ds, RowScratch& scratch) {
    if (ds.name != "update") throw DatasetError("An update dataset is expected, got '" + ds.name + "'!\n");
    UpdateData u{};
    u.n_scenarios = ds.batch_size;
    for (auto const& b : ds.buffers) {
        if (b.total_elements == 0) continue;
        if (ignored_by_power_flow(b.component)) continue;
        ComponentBuffer const cb = buffer_of(b, scratch);
        if (b.component == "line") u.line = cb;
        else if (b.component == "asym_line") u.asym_line = cb;
        else if (b.component == "generic_branch") u.generic_branch = cb;
        else if (b.component == "link") u.link = cb;
        else if (b.component == "three_winding_transformer") u.three_winding_transformer = cb;
        else if (b.component == "transformer_tap_regulator") u.transformer_tap_regulator = cb;
        else if (b.component == "transformer") u.transformer = cb;
        else if (b.component == "shunt") u.shunt = cb;
        else if (b.component == "source") u.source = cb;
        else if (b.component == "sym_gen") u.sym_gen = cb;
        else if (b.component == "asym_gen") u.asym_gen = cb;
        else if (b.component == "sym_load") u.sym_load = cb;
        else if (b.component == "asym_load") u.asym_load = cb;
        else if (b.component == "voltage_regulator") u.voltage_regulator = cb;
        else if (ignored_by_power_flow(b.component)) continue;
        else throw InvalidArgument("component '" + b.component + "' cannot be updated by pgm_b200\n");
    }
    return u;
}


// BatchCalculationError (job_dispatch.hpp:208-224): combined message, failed scenario numbers, one message per scenario
struct ApiBatchFailure {
    std::string message;
    std::vector<PGM_Idx> scenarios;
    std::vector<std::string> errors;
};

void set_output_slot(OutputData& od, std::string const& c, void* rows) {
    if (c == "node") od.node = rows;
    else if (c == "line") od.line = rows;
    else if (c == "asym_line") od.asym_line = rows;
    else if (c == "generic_branch") od.generic_branch = rows;
    else if (c == "link") od.link = rows;
    else if (c == "three_winding_transformer") od.three_winding_transformer = rows;
    else if (c == "transformer_tap_regulator") od.transformer_tap_regulator = rows;
    else if (c == "transformer") od.transformer = rows;
    else if (c == "shunt") od.shunt = rows;
    else if (c == "source") od.source = rows;
    else if (c == "sym_gen") od.sym_gen = rows;
    else if (c == "asym_gen") od.asym_gen = rows;
    else if (c == "sym_load") od.sym_load = rows;
    else if (c == "asym_load") od.asym_load = rows;
    else if (c == "voltage_regulator") od.voltage_regulator = rows;
}

// calculate_single_batch_dimension_impl (power_grid_model_c/src/model.cpp:188-204): one engine call for the whole batch
void calculate_single_dimension(Model& m, ModelOptions const& mo, Dataset const& out_ds, Dataset const* batch) {
    if (batch != nullptr && (!batch->is_batch || !out_ds.is_batch)) {
        throw CalculationError("If batch_dataset is provided. Both batch_dataset and output_dataset should be a batch!\n");
    }
    PGM_Idx const n_scn = batch != nullptr ? batch->batch_size : 1;
    if (out_ds.batch_size != n_scn) throw DatasetError("The batch sizes of the update and the output dataset differ!\n");
    if (n_scn == 0) return; // empty batch: nothing to calculate (job_dispatch.hpp:43-46)

    OutputData od{};
    RowScratch scratch;
    for (auto const& b : out_ds.buffers) {
        if (b.total_elements == 0) continue;
        Idx const count = m.component_count(b.component);
        if (count < 0) throw InvalidArgument("component '" + b.component + "' has no power-flow output in pgm_b200\n");
        if (b.elements_per_scenario != count) {
            throw DatasetError("The output buffer of '" + b.component + "' must hold exactly the model's " + std::to_string(count) + " elements per scenario!\n");
        }
        set_output_slot(od, b.component, scratch.rows_out(b));
    }
    std::vector<int32_t> status(static_cast<size_t>(n_scn), 0);
    int64_t failed;
    if (batch != nullptr) {
        UpdateData const ud = update_of(*batch, scratch);
        failed = m.calculate(mo, &ud, od, nullptr, status.data());
    } else {
        failed = m.calculate(mo, nullptr, od, nullptr, status.data());
    }
    scratch.scatter_outputs();
    if (failed == 0) return;
    std::string const& all = m.batch_message;
    if (batch == nullptr) {
        // a single calculation reports the solver's exception itself (PGM_regular_error)
        std::string msg = all;
        auto const colon = msg.find(": ");
        if (msg.rfind("Error in batch #", 0) == 0 && colon != std::string::npos) msg = msg.substr(colon + 2);
        throw CalculationError(msg);
    }
    ApiBatchFailure f;
    f.message = all;
    std::string const tag = "Error in batch #";
    size_t pos = all.find(tag);
    while (pos != std::string::npos) {
        size_t const next = all.find(tag, pos + tag.size());
        std::string const entry = all.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
        size_t const colon = entry.find(": ");
        PGM_Idx const scenario = std::strtoll(entry.c_str() + tag.size(), nullptr, 10);
        std::string msg = colon == std::string::npos ? entry : entry.substr(colon + 2);
        if (!msg.empty() && msg.back() == '\n') msg.pop_back(); // the line break the batch message adds behind what()
        f.scenarios.push_back(scenario);
        f.errors.push_back(std::move(msg));
        pos = next;
    }
    if (f.scenarios.empty()) { // messages were not itemised: fall back on the status array
        for (size_t s = 0; s != status.size(); ++s) {
            if (status[s] != 0) {
                f.scenarios.push_back(static_cast<PGM_Idx>(s));
                f.errors.emplace_back("scenario failed\n");
            }
        }
    }
    throw f;
}

// calculate_multi_dimensional_impl (model.cpp:290-334): a linked list of batch datasets is a cartesian product; every scenario of
// an outer dimension is applied as a permanent update to a copy of the model, the inner dimensions run on the copy and write the
// matching slice of the output.  The innermost dimension is the batch the GPU pipeline takes in one call.
void calculate_multi_dimensional(Model& m, ModelOptions const& mo, Dataset const& out_ds, Dataset const* batch) {
    if (batch == nullptr || batch->next == nullptr) {
        calculate_single_dimension(m, mo, out_ds, batch);
        return;
    }
    PGM_Idx stride = 1;
    for (Dataset const* d = batch->next; d != nullptr; d = d->next) stride *= d->batch_size;
    ApiBatchFailure all;
    for (PGM_Idx i = 0; i != batch->batch_size; ++i) {
        auto fail_all = [&](std::string const& what) { // a failure outside the inner batch fails every scenario of the slice
            all.message = what;
            for (PGM_Idx k = 0; k != stride; ++k) {
                all.scenarios.push_back(i * stride + k); // the reference lists 0..stride-1 here (model.cpp:253); the slice's own
                all.errors.push_back(what);              // scenario numbers are what the caller can act on
            }
        };
        try {
            Dataset const single = batch->individual_scenario(i);
            Dataset const slice = out_ds.slice_scenarios(i * stride, (i + 1) * stride);
            std::unique_ptr<Model> local = m.clone();
            RowScratch scratch;
            local->update_permanent(update_of(single, scratch));
            calculate_multi_dimensional(*local, mo, slice, batch->next);
        } catch (ApiBatchFailure const& f) {
            all.message = f.message;
            for (PGM_Idx s : f.scenarios) all.scenarios.push_back(s + i * stride);
            all.errors.insert(all.errors.end(), f.errors.begin(), f.errors.end());
        } catch (std::exception const& e) {
            fail_all(e.what());
        }
    }
    if (!all.scenarios.empty()) throw all;
}

} // namespace

struct PGM_PowerGridModel {
    std::unique_ptr<Model> model;
};

extern "C" {

// ---- handle --------------------------------------------------------------------------------------------------------
PGM_Handle* PGM_create_handle(void) {
    try {
        return new PGM_Handle{};
    } catch (...) {
        return nullptr;
    }
}
void PGM_destroy_handle(PGM_Handle* handle) { delete handle; }
PGM_Idx PGM_error_code(PGM_Handle const* handle) { return handle != nullptr ? handle->err_code : 0; }
char const* PGM_error_message(PGM_Handle const* handle) { return handle != nullptr ? handle->err_msg.c_str() : nullptr; }
PGM_Idx PGM_n_failed_scenarios(PGM_Handle const* handle) {
    return handle != nullptr ? static_cast<PGM_Idx>(handle->failed_scenarios.size()) : 0;
}
PGM_Idx const* PGM_failed_scenarios(PGM_Handle const* handle) {
    return handle != nullptr ? handle->failed_scenarios.data() : nullptr;
}
char const** PGM_batch_errors(PGM_Handle const* handle) {
    if (handle == nullptr) return nullptr;
    handle->batch_errs_c_str.clear();
    for (auto const& s : handle->batch_errs) handle->batch_errs_c_str.push_back(s.c_str());
    return handle->batch_errs_c_str.data();
}
void PGM_clear_error(PGM_Handle* handle) { clear(handle); }
char const* PGM_version(void) { return "pgm_b200 0.1 (power-flow path of power-grid-model on sm_100a)"; }

// ---- options -------------------------------------------------------------------------------------------------------
PGM_Options* PGM_create_options(PGM_Handle* handle) {
    return call(handle, [] { return new PGM_Options{}; });
}
void PGM_destroy_options(PGM_Options* opt) { delete opt; }
void PGM_set_calculation_type(PGM_Handle* handle, PGM_Options* opt, PGM_Idx type) {
    call(handle, [&] { deref(opt).calculation_type = type; });
}
void PGM_set_calculation_method(PGM_Handle* handle, PGM_Options* opt, PGM_Idx method) {
    call(handle, [&] { deref(opt).calculation_method = method; });
}
void PGM_set_symmetric(PGM_Handle* handle, PGM_Options* opt, PGM_Idx sym) {
    call(handle, [&] { deref(opt).symmetric = sym; });
}
void PGM_set_err_tol(PGM_Handle* handle, PGM_Options* opt, double err_tol) {
    call(handle, [&] { deref(opt).err_tol = err_tol; });
}
void PGM_set_max_iter(PGM_Handle* handle, PGM_Options* opt, PGM_Idx max_iter) {
    call(handle, [&] { deref(opt).max_iter = max_iter; });
}
void PGM_set_threading(PGM_Handle* handle, PGM_Options* opt, PGM_Idx threading) {
    call(handle, [&] { deref(opt).threading = threading; });
}
void PGM_set_short_circuit_voltage_scaling(PGM_Handle* handle, PGM_Options* opt, PGM_Idx short_circuit_voltage_scaling) {
    call(handle, [&] { deref(opt).short_circuit_voltage_scaling = short_circuit_voltage_scaling; });
}
void PGM_set_tap_changing_strategy(PGM_Handle* handle, PGM_Options* opt, PGM_Idx tap_changing_strategy) {
    call(handle, [&] { deref(opt).tap_changing_strategy = tap_changing_strategy; });
}
void PGM_set_experimental_features(PGM_Handle* handle, PGM_Options* opt, PGM_Idx experimental_features) {
    call(handle, [&] { deref(opt).experimental_features = experimental_features; });
}

// ---- datasets (row-based) ------------------------------------------------------------------------------------------
PGM_ConstDataset* PGM_create_dataset_const(PGM_Handle* handle, char const* dataset, PGM_Idx is_batch, PGM_Idx batch_size) {
    return call(handle, [&] { return new PGM_ConstDataset{dataset, is_batch, batch_size}; });
}
PGM_ConstDataset* PGM_create_dataset_const_from_mutable(PGM_Handle* handle, PGM_MutableDataset const* mutable_dataset) {
    return call(handle, [&] {
        Dataset const& src = deref(mutable_dataset);
        auto* ds = new PGM_ConstDataset{src.name.c_str(), src.is_batch ? 1 : 0, src.batch_size};
        ds->buffers = src.buffers;
        return ds;
    });
}
void PGM_destroy_dataset_const(PGM_ConstDataset* dataset) { delete dataset; }
void PGM_dataset_const_add_buffer(PGM_Handle* handle, PGM_ConstDataset* dataset, char const* component,
                                  PGM_Idx elements_per_scenario, PGM_Idx total_elements, PGM_Idx const* indptr,
                                  void const* data) {
    call(handle, [&] {
        deref(dataset).add_buffer(component, elements_per_scenario, total_elements, indptr, const_cast<void*>(data), true);
    });
}
void PGM_dataset_const_add_attribute_buffer(PGM_Handle* handle, PGM_ConstDataset* dataset, char const* component,
                                            char const* attribute, void const* data) {
    call(handle, [&] { deref(dataset).add_attribute_buffer(component, attribute, const_cast<void*>(data)); });
}
void PGM_dataset_const_set_next_cartesian_product_dimension(PGM_Handle* handle, PGM_ConstDataset* dataset,
                                                            PGM_ConstDataset const* next_dataset) {
    call(handle, [&] { deref(dataset).set_next(next_dataset); });
}
PGM_MutableDataset* PGM_create_dataset_mutable(PGM_Handle* handle, char const* dataset, PGM_Idx is_batch, PGM_Idx batch_size) {
    return call(handle, [&] { return new PGM_MutableDataset{dataset, is_batch, batch_size}; });
}
void PGM_destroy_dataset_mutable(PGM_MutableDataset* dataset) { delete dataset; }
void PGM_dataset_mutable_add_buffer(PGM_Handle* handle, PGM_MutableDataset* dataset, char const* component,
                                    PGM_Idx elements_per_scenario, PGM_Idx total_elements, PGM_Idx const* indptr, void* data) {
    call(handle, [&] { deref(dataset).add_buffer(component, elements_per_scenario, total_elements, indptr, data, false); });
}
void PGM_dataset_mutable_add_attribute_buffer(PGM_Handle* handle, PGM_MutableDataset* dataset, char const* component,
                                              char const* attribute, void* data) {
    call(handle, [&] { deref(dataset).add_attribute_buffer(component, attribute, data); });
}
PGM_DatasetInfo const* PGM_dataset_const_get_info(PGM_Handle* handle, PGM_ConstDataset const* dataset) {
    return call(handle, [&] { return reinterpret_cast<PGM_DatasetInfo const*>(static_cast<Dataset const*>(&deref(dataset))); });
}
PGM_DatasetInfo const* PGM_dataset_mutable_get_info(PGM_Handle* handle, PGM_MutableDataset const* dataset) {
    return call(handle, [&] { return reinterpret_cast<PGM_DatasetInfo const*>(static_cast<Dataset const*>(&deref(dataset))); });
}

// ---- dataset info (dataset.h:27-138) -----------------------------------------------------------------------------
#define PGMB_INFO(info) (*reinterpret_cast<Dataset const*>(&deref(info)))
char const* PGM_dataset_info_name(PGM_Handle* handle, PGM_DatasetInfo const* info) {
    return call(handle, [&] { return PGMB_INFO(info).name.c_str(); });
}
PGM_Idx PGM_dataset_info_is_batch(PGM_Handle* handle, PGM_DatasetInfo const* info) {
    return call(handle, [&] { return static_cast<PGM_Idx>(PGMB_INFO(info).is_batch); });
}
PGM_Idx PGM_dataset_info_batch_size(PGM_Handle* handle, PGM_DatasetInfo const* info) {
    return call(handle, [&] { return PGMB_INFO(info).batch_size; });
}
PGM_Idx PGM_dataset_info_n_components(PGM_Handle* handle, PGM_DatasetInfo const* info) {
    return call(handle, [&] { return static_cast<PGM_Idx>(PGMB_INFO(info).buffers.size()); });
}
char const* PGM_dataset_info_component_name(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx) {
    return call(handle, [&] { return PGMB_INFO(info).at(component_idx).component.c_str(); });
}
PGM_Idx PGM_dataset_info_elements_per_scenario(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx) {
    return call(handle, [&] { return PGMB_INFO(info).at(component_idx).elements_per_scenario; });
}
PGM_Idx PGM_dataset_info_total_elements(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx) {
    return call(handle, [&] { return PGMB_INFO(info).at(component_idx).total_elements; });
}
PGM_Idx PGM_dataset_info_has_attribute_indications(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx) {
    return call(handle, [&] {
        auto const& b = PGMB_INFO(info).at(component_idx);
        return static_cast<PGM_Idx>(b.has_indications || (b.columnar() && !b.attributes.empty()));
    });
}
PGM_Idx PGM_dataset_info_n_attribute_indications(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx) {
    return call(handle, [&] {
        auto const& b = PGMB_INFO(info).at(component_idx);
        return static_cast<PGM_Idx>(b.has_indications ? b.indications.size() : b.attributes.size());
    });
}
char const* PGM_dataset_info_attribute_name(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx,
                                            PGM_Idx attribute_idx) {
    return call(handle, [&] {
        auto const& b = PGMB_INFO(info).at(component_idx);
        PGM_Idx const n = static_cast<PGM_Idx>(b.has_indications ? b.indications.size() : b.attributes.size());
        if (attribute_idx < 0 || attribute_idx >= n) throw std::out_of_range("Index out of range!\n");
        return b.has_indications ? b.indications[static_cast<size_t>(attribute_idx)]->name
                                 : b.attributes[static_cast<size_t>(attribute_idx)].attribute->name;
    });
}
#undef PGMB_INFO

// ---- model ---------------------------------------------------------------------------------------------------------
PGM_PowerGridModel* PGM_create_model(PGM_Handle* handle, double system_frequency, PGM_ConstDataset const* input_dataset) {
    return call(handle, [&] {
        Dataset const& ds = deref(input_dataset);
        if (ds.name != "input" || ds.is_batch) throw DatasetError("PGM_create_model takes a single (non-batch) input dataset!\n");
        InputData in{};
        RowScratch scratch;
        for (auto const& b : ds.buffers) {
            if (b.total_elements == 0 || ignored_by_power_flow(b.component)) continue;
            ComponentBuffer const cb{b.total_elements, nullptr, scratch.rows_in(b)};
            if (b.component == "node") in.node = cb;
            else if (b.component == "line") in.line = cb;
            else if (b.component == "asym_line") in.asym_line = cb;
            else if (b.component == "generic_branch") in.generic_branch = cb;
            else if (b.component == "link") in.link = cb;
            else if (b.component == "three_winding_transformer") in.three_winding_transformer = cb;
            else if (b.component == "transformer_tap_regulator") in.transformer_tap_regulator = cb;
            else if (b.component == "transformer") in.transformer = cb;
            else if (b.component == "shunt") in.shunt = cb;
            else if (b.component == "source") in.source = cb;
            else if (b.component == "sym_gen") in.sym_gen = cb;
            else if (b.component == "asym_gen") in.asym_gen = cb;
            else if (b.component == "sym_load") in.sym_load = cb;
            else if (b.component == "asym_load") in.asym_load = cb;
            else if (b.component == "voltage_regulator") in.voltage_regulator = cb;
            else throw InvalidArgument("component '" + b.component + "' is not built by pgm_b200 (power-flow path only)\n");
        }
        auto m = std::make_unique<PGM_PowerGridModel>();
        m->model = std::make_unique<Model>(system_frequency, in);
        return m.release();
    });
}

void PGM_update_model(PGM_Handle* handle, PGM_PowerGridModel* model, PGM_ConstDataset const* update_dataset) {
    call(handle, [&] {
        Dataset const& ds = deref(update_dataset);
        if (ds.is_batch) throw DatasetError("PGM_update_model takes a single (non-batch) update dataset!\n");
        RowScratch scratch;
        deref(model).model->update_permanent(update_of(ds, scratch));
    });
}

PGM_PowerGridModel* PGM_copy_model(PGM_Handle* handle, PGM_PowerGridModel const* model) {
    return call(handle, [&] {
        auto m = std::make_unique<PGM_PowerGridModel>();
        m->model = deref(model).model->clone();
        return m.release();
    });
}

void PGM_get_indexer(PGM_Handle* handle, PGM_PowerGridModel const* model, char const* component, PGM_Idx size,
                     PGM_ID const* ids, PGM_Idx* indexer) {
    call(handle, [&] {
        if (component == nullptr || (size != 0 && (ids == nullptr || indexer == nullptr))) {
            throw InvalidArgument("Received null pointer where a valid pointer was expected.\n");
        }
        deref(model).model->get_indexer(component, ids, size, indexer);
    });
}

void PGM_calculate(PGM_Handle* handle, PGM_PowerGridModel* model, PGM_Options const* opt,
                   PGM_MutableDataset const* output_dataset, PGM_ConstDataset const* batch_dataset) {
    ApiBatchFailure batch_failure;
    bool batch_failed = false;
    call(handle, [&] {
        Model& m = *deref(model).model;
        PGM_Options const& o = deref(opt);
        Dataset const& out_ds = deref(output_dataset);
        if (o.calculation_type < PGM_power_flow || o.calculation_type > PGM_short_circuit) {
            throw InvalidArgument("CalculationType is not implemented for #" + std::to_string(o.calculation_type) + "!\n");
        }
        if (o.calculation_type != PGM_power_flow) {
            throw CalculationError("pgm_b200 provides calculation type power_flow only (state estimation and short circuit are "
                                   "outside the accelerated path)\n");
        }
        // automatic tap changing (optimizer/tap_position_optimizer.hpp) acts on the transformer_tap_regulator components of the
        // model; with none in the model every valid strategy is the plain power flow
        if (o.tap_changing_strategy < 0 || o.tap_changing_strategy > 4) {
            throw InvalidArgument("get_optimizer_type is not implemented for #" + std::to_string(o.tap_changing_strategy) + "!\n");
        }
        if (o.symmetric != PGM_symmetric && o.symmetric != PGM_asymmetric) throw InvalidArgument("get_calculation_symmetry is not implemented for this value\n");
        int32_t method;
        switch (o.calculation_method) {
        case PGM_default_method:
        case PGM_newton_raphson: method = PGMB_METHOD_NEWTON_RAPHSON; break;
        case PGM_linear: method = PGMB_METHOD_LINEAR; break;
        case PGM_iterative_current: method = PGMB_METHOD_ITERATIVE_CURRENT; break;
        case PGM_linear_current: method = PGMB_METHOD_LINEAR_CURRENT; break;
        default:
            throw CalculationError("The calculation method is invalid for this calculation!\n"); // InvalidCalculationMethod
        }
        bool const sym = o.symmetric == PGM_symmetric;
        if (out_ds.name != (sym ? "sym_output" : "asym_output")) {
            throw DatasetError("The output dataset '" + out_ds.name + "' does not match the calculation symmetry!\n");
        }
        if (o.max_iter < 0 || o.max_iter > (PGM_Idx{1} << 30)) throw InvalidArgument("max_iter out of range\n");
        // threading (job_dispatch.hpp:166-171): < 0 = sequential, 0 = hardware concurrency, n = n threads.  It only matters for
        // batches that take the per-scenario route (own topology per scenario); load batches run on the GPU in one pipeline.
        int32_t const threading = o.threading < 0 ? 1 : static_cast<int32_t>(std::min<PGM_Idx>(o.threading, 1 << 20));
        ModelOptions mo{method, sym, o.err_tol, o.max_iter, device_ordinal(), threading};
        mo.tap_strategy = static_cast<int32_t>(o.tap_changing_strategy);
        // the output of a cartesian product holds the product of all dimension sizes (model.cpp:290-334 asserts it)
        if (batch_dataset != nullptr) {
            PGM_Idx total = 1;
            for (Dataset const* d = batch_dataset; d != nullptr; d = d->next) {
                total *= d->batch_size;
            }
            if (out_ds.batch_size != total) {
                throw DatasetError("The batch size of the output dataset (" + std::to_string(out_ds.batch_size) +
                                   ") does not match the product of the batch dimensions (" + std::to_string(total) + ")!\n");
            }
        }
        try {
            calculate_multi_dimensional(m, mo, out_ds, batch_dataset);
        } catch (ApiBatchFailure& f) {
            batch_failure = std::move(f);
            batch_failed = true;
        }
    });
    if (!batch_failed || handle == nullptr) return;
    handle->err_code = PGM_batch_error;
    handle->err_msg = std::move(batch_failure.message);
    handle->failed_scenarios = std::move(batch_failure.scenarios);
    handle->batch_errs = std::move(batch_failure.errors);
}

void PGM_destroy_model(PGM_PowerGridModel* model) { delete model; }

} // extern "C"
