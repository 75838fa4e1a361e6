// temporary: model-level entry points not implemented yet
#include "capi_common.hpp"
using namespace pgmb;
#define NOT_IMPL(name) return guarded([] { throw InvalidArgument(name " is not implemented yet"); })
extern "C" {
int pgmb_model_create(double, const pgmb_input_data*, pgmb_model**) { NOT_IMPL("pgmb_model_create"); }
void pgmb_model_destroy(pgmb_model*) {}
int pgmb_model_update(pgmb_model*, const pgmb_update_data*) { NOT_IMPL("pgmb_model_update"); }
int pgmb_model_calculate(pgmb_model*, const pgmb_options*, const pgmb_update_data*, const pgmb_output_data*, int32_t*, int32_t*) { NOT_IMPL("pgmb_model_calculate"); }
int pgmb_model_get_index(pgmb_model*, int64_t, const char*, const int64_t**, int64_t*) { NOT_IMPL("pgmb_model_get_index"); }
int pgmb_model_get_real(pgmb_model*, int64_t, int32_t, const char*, const double**, int64_t*) { NOT_IMPL("pgmb_model_get_real"); }
int64_t pgmb_model_n_math_groups(pgmb_model*) { return 0; }
int pgmb_model_last_timing(pgmb_model*, double*) { NOT_IMPL("pgmb_model_last_timing"); }
int pgmb_fictional_grid_create(const pgmb_grid_option*, uint32_t, pgmb_fictional_grid**) { NOT_IMPL("pgmb_fictional_grid_create"); }
void pgmb_fictional_grid_destroy(pgmb_fictional_grid*) {}
int pgmb_fictional_grid_get(pgmb_fictional_grid*, const char*, const void**, int64_t*) { NOT_IMPL("pgmb_fictional_grid_get"); }
int pgmb_fictional_grid_batch(pgmb_fictional_grid*, int64_t, uint32_t, void*, void*) { NOT_IMPL("pgmb_fictional_grid_batch"); }
}
