// C-ABI, model level (include/pgm_b200.h): PGM_create_model / PGM_update_model / PGM_calculate analogues.
#include "capi_common.hpp"
#include "model.hpp"

using namespace pgmb;

struct pgmb_model {
    std::unique_ptr<Model> model;
};

namespace {
ComponentBuffer cb(pgmb_component_buffer const& b) { return {b.n, b.indptr, b.data}; }
InputData input_of(pgmb_input_data const& in) {
    return {cb(in.node), cb(in.line), cb(in.transformer), cb(in.shunt), cb(in.source), cb(in.sym_gen), cb(in.asym_gen),
            cb(in.sym_load), cb(in.asym_load), cb(in.voltage_regulator), cb(in.asym_line), cb(in.generic_branch), cb(in.link),
            cb(in.three_winding_transformer), cb(in.transformer_tap_regulator)};
}
UpdateData update_of(pgmb_update_data const& u) {
    return {u.n_scenarios, cb(u.line), cb(u.transformer), cb(u.shunt), cb(u.source), cb(u.sym_gen), cb(u.asym_gen),
            cb(u.sym_load), cb(u.asym_load), cb(u.voltage_regulator), cb(u.asym_line), cb(u.generic_branch), cb(u.link),
            cb(u.three_winding_transformer), cb(u.transformer_tap_regulator)};
}
} // namespace

extern "C" {

int pgmb_model_create(double system_frequency, const pgmb_input_data* input, pgmb_model** out) {
    return guarded([&] {
        if (input == nullptr || out == nullptr) throw InvalidArgument("null argument");
        auto h = std::make_unique<pgmb_model>();
        h->model = std::make_unique<Model>(system_frequency, input_of(*input));
        *out = h.release();
    });
}
void pgmb_model_destroy(pgmb_model* model) { delete model; }

int pgmb_model_update(pgmb_model* model, const pgmb_update_data* update) {
    return guarded([&] {
        if (model == nullptr || update == nullptr) throw InvalidArgument("null argument");
        model->model->update_permanent(update_of(*update));
    });
}

int pgmb_model_calculate(pgmb_model* model, const pgmb_options* opt, const pgmb_update_data* update,
                         const pgmb_output_data* output, int32_t* n_iter, int32_t* status) {
    int64_t failed = 0;
    int const rc = guarded([&] {
        if (model == nullptr || opt == nullptr || output == nullptr) throw InvalidArgument("null argument");
        if (opt->max_iter < 0 || opt->max_iter > (int64_t{1} << 30)) throw InvalidArgument("max_iter out of range");
        ModelOptions const mo{opt->calculation_method, opt->symmetric != 0, opt->err_tol, opt->max_iter, opt->first_device, opt->threading,
                              opt->n_devices, opt->flags, opt->tap_changing_strategy};
        if (mo.tap_strategy < 0 || mo.tap_strategy > 4) throw InvalidArgument("tap_changing_strategy out of range");
        OutputData const od{output->node, output->line, output->transformer, output->shunt, output->source,
                            output->sym_gen, output->asym_gen, output->sym_load, output->asym_load,
                            output->voltage_regulator, output->asym_line, output->generic_branch, output->link,
                            output->three_winding_transformer, output->transformer_tap_regulator};
        if (update != nullptr) {
            UpdateData const ud = update_of(*update);
            failed = model->model->calculate(mo, &ud, od, n_iter, status);
        } else {
            failed = model->model->calculate(mo, nullptr, od, n_iter, status);
        }
    });
    if (rc != PGMB_OK) return rc;
    if (failed != 0) {
        g_last_error = model->model->batch_message;
        return PGMB_ERR_BATCH;
    }
    return PGMB_OK;
}

int64_t pgmb_model_n_math_groups(pgmb_model* model) {
    int64_t n = -1;
    guarded([&] {
        if (model == nullptr) throw InvalidArgument("null argument");
        n = model->model->n_math_groups();
    });
    return n;
}

int pgmb_model_batch_pf_input(pgmb_model* model, const pgmb_update_data* update, int32_t symmetric, int64_t math_group,
                              double* s_injection, double* source_u_ref) {
    return guarded([&] {
        if (model == nullptr || update == nullptr || s_injection == nullptr || source_u_ref == nullptr) throw InvalidArgument("null argument");
        model->model->batch_pf_input(update_of(*update), symmetric != 0, math_group, s_injection, source_u_ref);
    });
}

int pgmb_model_outage_plan(pgmb_model* model, const pgmb_update_data* update, int32_t symmetric, int64_t* plan) {
    return guarded([&] {
        if (model == nullptr || update == nullptr || plan == nullptr) throw InvalidArgument("null argument");
        model->model->outage_plan_summary(update_of(*update), symmetric != 0, plan);
    });
}

int pgmb_model_get_index(pgmb_model* model, int64_t math_group, const char* name, const int64_t** data, int64_t* size) {
    return guarded([&] {
        if (model == nullptr || name == nullptr) throw InvalidArgument("null argument");
        auto const& v = model->model->get_index(math_group, name);
        *data = v.data();
        *size = static_cast<int64_t>(v.size());
    });
}

int pgmb_model_get_real(pgmb_model* model, int64_t math_group, int32_t symmetric, const char* name, const double** data,
                        int64_t* size) {
    return guarded([&] {
        if (model == nullptr || name == nullptr) throw InvalidArgument("null argument");
        auto const& v = model->model->get_real(math_group, symmetric != 0, name);
        *data = v.data();
        *size = static_cast<int64_t>(v.size());
    });
}

int pgmb_model_last_timing(pgmb_model* model, double* ms6) {
    return guarded([&] {
        if (model == nullptr || ms6 == nullptr) throw InvalidArgument("null argument");
        for (int i = 0; i != 6; ++i) ms6[i] = model->model->timing[i];
    });
}

int pgmb_model_device_pipeline_ms(pgmb_model* model, double* ms) {
    return guarded([&] {
        if (model == nullptr || ms == nullptr) throw InvalidArgument("null argument");
        *ms = model->model->timing[6];
    });
}

} // extern "C"
