/* The reference's own C API names for the power-flow path, exported by libpgm_b200.so.
 *
 * A C client (or the reference's ctypes wrapper restricted to row-based buffers) that drives power flow through
 * `PGM_create_handle / PGM_create_options / PGM_create_dataset_* / PGM_create_model / PGM_calculate` links against this library
 * unchanged: same symbol names, argument meaning and error behaviour as
 *   power_grid_model_c/include/power_grid_model_c/handle.h:32-115, options.h:38-138, dataset.h:140-320, model.h:33-125
 * of the reference.  Scope = what the engine builds:
 *   - calculation type power_flow; methods default / newton_raphson / linear / iterative_current / linear_current; symmetric and
 *     asymmetric; single and batch;
 *   - row-based and columnar ("attribute") buffers, dense or sparse; a cartesian product of update datasets
 *     (PGM_dataset_const_set_next_cartesian_product_dimension): outer dimensions are applied as permanent updates to a model copy,
 *     the innermost dimension is the batch one GPU call solves;
 *   - the meta-data tables (PGM_meta_*) of every dataset and component of the reference, PGM_create_buffer / PGM_buffer_* and
 *     the dataset info calls, so a client sizes and fills its buffers the way the reference's wrapper does;
 *   - components: node, line, asym_line, link, generic_branch, transformer, three_winding_transformer, shunt, source, sym_gen,
 *     asym_gen, sym_load, asym_load, voltage_regulator, transformer_tap_regulator; sensors and faults may be present in the input
 *     dataset and are ignored by power flow like in the reference;
 *   - tap_changing_strategy: the automatic tap changer of the reference (optimizer/tap_position_optimizer.hpp) around the power
 *     flows of every scenario: any_valid_tap, min_voltage_tap, max_voltage_tap, fast_any_tap;
 *   - JSON and msgpack (de)serialization of datasets and the writable datasets of the deserializer (serialization.h, dataset.h).
 *   - the PGM_def_* pointer constants of dataset_definitions.h (include/pgm_b200_dataset_definitions.h, 859 symbols).
 * Not provided: calculation types other than power flow (state estimation, short circuit).
 * There is no CPU fallback: PGM_calculate on a host without a CUDA device reports PGM_regular_error.
 */
#ifndef PGM_B200_CAPI_H
#define PGM_B200_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGM_API __attribute__((visibility("default")))

typedef int64_t PGM_Idx;
typedef int32_t PGM_ID;
typedef struct PGM_Handle PGM_Handle;
typedef struct PGM_Options PGM_Options;
typedef struct PGM_ConstDataset PGM_ConstDataset;
typedef struct PGM_MutableDataset PGM_MutableDataset;
typedef struct PGM_PowerGridModel PGM_PowerGridModel;
typedef struct PGM_MetaDataset PGM_MetaDataset;
typedef struct PGM_MetaComponent PGM_MetaComponent;
typedef struct PGM_MetaAttribute PGM_MetaAttribute;
typedef struct PGM_DatasetInfo PGM_DatasetInfo;
typedef struct PGM_WritableDataset PGM_WritableDataset;
typedef struct PGM_Serializer PGM_Serializer;
typedef struct PGM_Deserializer PGM_Deserializer;

enum PGM_CalculationType { PGM_power_flow = 0, PGM_state_estimation = 1, PGM_short_circuit = 2 };
enum PGM_CalculationMethod {
    PGM_default_method = -128,
    PGM_linear = 0,
    PGM_newton_raphson = 1,
    PGM_iterative_linear = 2,
    PGM_iterative_current = 3,
    PGM_linear_current = 4,
    PGM_iec60909 = 5
};
enum PGM_SymmetryType { PGM_asymmetric = 0, PGM_symmetric = 1 };
enum PGM_CType { PGM_int32 = 0, PGM_int8 = 1, PGM_double = 2, PGM_double3 = 3 };
enum PGM_ErrorCode { PGM_no_error = 0, PGM_regular_error = 1, PGM_batch_error = 2, PGM_serialization_error = 3 };

/* handle.h */
PGM_API PGM_Handle* PGM_create_handle(void);
PGM_API void PGM_destroy_handle(PGM_Handle* handle);
PGM_API PGM_Idx PGM_error_code(PGM_Handle const* handle);
PGM_API char const* PGM_error_message(PGM_Handle const* handle);
PGM_API PGM_Idx PGM_n_failed_scenarios(PGM_Handle const* handle);
PGM_API PGM_Idx const* PGM_failed_scenarios(PGM_Handle const* handle);
PGM_API char const** PGM_batch_errors(PGM_Handle const* handle);
PGM_API void PGM_clear_error(PGM_Handle* handle);
PGM_API char const* PGM_version(void);

/* options.h */
PGM_API PGM_Options* PGM_create_options(PGM_Handle* handle);
PGM_API void PGM_destroy_options(PGM_Options* opt);
PGM_API void PGM_set_calculation_type(PGM_Handle* handle, PGM_Options* opt, PGM_Idx type);
PGM_API void PGM_set_calculation_method(PGM_Handle* handle, PGM_Options* opt, PGM_Idx method);
PGM_API void PGM_set_symmetric(PGM_Handle* handle, PGM_Options* opt, PGM_Idx sym);
PGM_API void PGM_set_err_tol(PGM_Handle* handle, PGM_Options* opt, double err_tol);
PGM_API void PGM_set_max_iter(PGM_Handle* handle, PGM_Options* opt, PGM_Idx max_iter);
PGM_API void PGM_set_threading(PGM_Handle* handle, PGM_Options* opt, PGM_Idx threading);
PGM_API void PGM_set_short_circuit_voltage_scaling(PGM_Handle* handle, PGM_Options* opt, PGM_Idx short_circuit_voltage_scaling);
PGM_API void PGM_set_tap_changing_strategy(PGM_Handle* handle, PGM_Options* opt, PGM_Idx tap_changing_strategy);
PGM_API void PGM_set_experimental_features(PGM_Handle* handle, PGM_Options* opt, PGM_Idx experimental_features);

/* meta_data.h:32-190 -- datasets input / update / sym_output / asym_output / sc_output, every component of the reference, attribute
 * names, ctypes and offsets generated from the reference's definition files (tools/gen_meta_table.py) */
PGM_API PGM_Idx PGM_meta_n_datasets(PGM_Handle* handle);
PGM_API PGM_MetaDataset const* PGM_meta_get_dataset_by_idx(PGM_Handle* handle, PGM_Idx idx);
PGM_API PGM_MetaDataset const* PGM_meta_get_dataset_by_name(PGM_Handle* handle, char const* dataset);
PGM_API char const* PGM_meta_dataset_name(PGM_Handle* handle, PGM_MetaDataset const* dataset);
PGM_API PGM_Idx PGM_meta_n_components(PGM_Handle* handle, PGM_MetaDataset const* dataset);
PGM_API PGM_MetaComponent const* PGM_meta_get_component_by_idx(PGM_Handle* handle, PGM_MetaDataset const* dataset, PGM_Idx idx);
PGM_API PGM_MetaComponent const* PGM_meta_get_component_by_name(PGM_Handle* handle, char const* dataset, char const* component);
PGM_API char const* PGM_meta_component_name(PGM_Handle* handle, PGM_MetaComponent const* component);
PGM_API size_t PGM_meta_component_size(PGM_Handle* handle, PGM_MetaComponent const* component);
PGM_API size_t PGM_meta_component_alignment(PGM_Handle* handle, PGM_MetaComponent const* component);
PGM_API PGM_Idx PGM_meta_n_attributes(PGM_Handle* handle, PGM_MetaComponent const* component);
PGM_API PGM_MetaAttribute const* PGM_meta_get_attribute_by_idx(PGM_Handle* handle, PGM_MetaComponent const* component, PGM_Idx idx);
PGM_API PGM_MetaAttribute const* PGM_meta_get_attribute_by_name(PGM_Handle* handle, char const* dataset, char const* component,
                                                                char const* attribute);
PGM_API char const* PGM_meta_attribute_name(PGM_Handle* handle, PGM_MetaAttribute const* attribute);
PGM_API PGM_Idx PGM_meta_attribute_ctype(PGM_Handle* handle, PGM_MetaAttribute const* attribute);
PGM_API size_t PGM_meta_attribute_offset(PGM_Handle* handle, PGM_MetaAttribute const* attribute);
PGM_API int PGM_is_little_endian(PGM_Handle* handle);

/* buffer.h:40-108 */
PGM_API void* PGM_create_buffer(PGM_Handle* handle, PGM_MetaComponent const* component, PGM_Idx size);
PGM_API void PGM_destroy_buffer(void* ptr);
/* extension (not in the reference): 1 when `ptr` came from PGM_create_buffer as page-locked memory (>= 4 KB with a CUDA device
 * present), which the device pipeline fills directly, chunk by chunk, while the solver runs */
PGM_API int PGM_b200_buffer_is_page_locked(void const* ptr);
PGM_API void PGM_buffer_set_nan(PGM_Handle* handle, PGM_MetaComponent const* component, void* ptr, PGM_Idx buffer_offset,
                                PGM_Idx size);
PGM_API void PGM_buffer_set_value(PGM_Handle* handle, PGM_MetaAttribute const* attribute, void* buffer_ptr, void const* src_ptr,
                                  PGM_Idx buffer_offset, PGM_Idx size, PGM_Idx src_stride);
PGM_API void PGM_buffer_get_value(PGM_Handle* handle, PGM_MetaAttribute const* attribute, void const* buffer_ptr, void* dest_ptr,
                                  PGM_Idx buffer_offset, PGM_Idx size, PGM_Idx dest_stride);

/* dataset.h:27-138 (info) */
PGM_API char const* PGM_dataset_info_name(PGM_Handle* handle, PGM_DatasetInfo const* info);
PGM_API PGM_Idx PGM_dataset_info_is_batch(PGM_Handle* handle, PGM_DatasetInfo const* info);
PGM_API PGM_Idx PGM_dataset_info_batch_size(PGM_Handle* handle, PGM_DatasetInfo const* info);
PGM_API PGM_Idx PGM_dataset_info_n_components(PGM_Handle* handle, PGM_DatasetInfo const* info);
PGM_API char const* PGM_dataset_info_component_name(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx);
PGM_API PGM_Idx PGM_dataset_info_elements_per_scenario(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx);
PGM_API PGM_Idx PGM_dataset_info_total_elements(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx);
PGM_API PGM_Idx PGM_dataset_info_has_attribute_indications(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx);
PGM_API PGM_Idx PGM_dataset_info_n_attribute_indications(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx);
PGM_API char const* PGM_dataset_info_attribute_name(PGM_Handle* handle, PGM_DatasetInfo const* info, PGM_Idx component_idx,
                                                    PGM_Idx attribute_idx);
PGM_API PGM_DatasetInfo const* PGM_dataset_const_get_info(PGM_Handle* handle, PGM_ConstDataset const* dataset);
PGM_API PGM_DatasetInfo const* PGM_dataset_mutable_get_info(PGM_Handle* handle, PGM_MutableDataset const* dataset);

/* dataset.h:140-332 (row-based and columnar buffers, dense or sparse) */
PGM_API PGM_ConstDataset* PGM_create_dataset_const(PGM_Handle* handle, char const* dataset, PGM_Idx is_batch, PGM_Idx batch_size);
PGM_API PGM_ConstDataset* PGM_create_dataset_const_from_mutable(PGM_Handle* handle, PGM_MutableDataset const* mutable_dataset);
PGM_API void PGM_destroy_dataset_const(PGM_ConstDataset* dataset);
PGM_API void PGM_dataset_const_add_buffer(PGM_Handle* handle, PGM_ConstDataset* dataset, char const* component,
                                          PGM_Idx elements_per_scenario, PGM_Idx total_elements, PGM_Idx const* indptr,
                                          void const* data);
PGM_API void PGM_dataset_const_add_attribute_buffer(PGM_Handle* handle, PGM_ConstDataset* dataset, char const* component,
                                                    char const* attribute, void const* data);
PGM_API void PGM_dataset_const_set_next_cartesian_product_dimension(PGM_Handle* handle, PGM_ConstDataset* dataset,
                                                                    PGM_ConstDataset const* next_dataset);
PGM_API PGM_MutableDataset* PGM_create_dataset_mutable(PGM_Handle* handle, char const* dataset, PGM_Idx is_batch,
                                                       PGM_Idx batch_size);
PGM_API void PGM_destroy_dataset_mutable(PGM_MutableDataset* dataset);
PGM_API void PGM_dataset_mutable_add_buffer(PGM_Handle* handle, PGM_MutableDataset* dataset, char const* component,
                                            PGM_Idx elements_per_scenario, PGM_Idx total_elements, PGM_Idx const* indptr,
                                            void* data);
PGM_API void PGM_dataset_mutable_add_attribute_buffer(PGM_Handle* handle, PGM_MutableDataset* dataset, char const* component,
                                                      char const* attribute, void* data);

/* model.h */
PGM_API PGM_PowerGridModel* PGM_create_model(PGM_Handle* handle, double system_frequency, PGM_ConstDataset const* input_dataset);
PGM_API void PGM_update_model(PGM_Handle* handle, PGM_PowerGridModel* model, PGM_ConstDataset const* update_dataset);
PGM_API PGM_PowerGridModel* PGM_copy_model(PGM_Handle* handle, PGM_PowerGridModel const* model);
PGM_API void PGM_get_indexer(PGM_Handle* handle, PGM_PowerGridModel const* model, char const* component, PGM_Idx size,
                             PGM_ID const* ids, PGM_Idx* indexer);
PGM_API void PGM_calculate(PGM_Handle* handle, PGM_PowerGridModel* model, PGM_Options const* opt,
                           PGM_MutableDataset const* output_dataset, PGM_ConstDataset const* batch_dataset);
PGM_API void PGM_destroy_model(PGM_PowerGridModel* model);

/* serialization.h:29-114 and the writable-dataset calls of dataset.h:241-269: JSON (serialization_format 0) and msgpack (1) in the
 * reference's dataset format (auxiliary/serialization/{deserializer,serializer}.hpp; csrc/capi_pgm_serialization.cpp).  Failures
 * are reported as PGM_serialization_error.  The deserializer owns its writable dataset; the caller supplies the buffers
 * (PGM_dataset_writable_set_buffer / _set_attribute_buffer) before PGM_deserializer_parse_to_buffer fills them. */
PGM_API PGM_Deserializer* PGM_create_deserializer_from_binary_buffer(PGM_Handle* handle, char const* data, PGM_Idx size,
                                                                     PGM_Idx serialization_format);
PGM_API PGM_Deserializer* PGM_create_deserializer_from_null_terminated_string(PGM_Handle* handle, char const* data_string,
                                                                              PGM_Idx serialization_format);
PGM_API PGM_WritableDataset* PGM_deserializer_get_dataset(PGM_Handle* handle, PGM_Deserializer* deserializer);
PGM_API void PGM_deserializer_parse_to_buffer(PGM_Handle* handle, PGM_Deserializer* deserializer);
PGM_API void PGM_destroy_deserializer(PGM_Deserializer* deserializer);
PGM_API PGM_Serializer* PGM_create_serializer(PGM_Handle* handle, PGM_ConstDataset const* dataset, PGM_Idx serialization_format);
PGM_API void PGM_serializer_get_to_binary_buffer(PGM_Handle* handle, PGM_Serializer* serializer, PGM_Idx use_compact_list,
                                                 char const** data, PGM_Idx* size);
PGM_API char const* PGM_serializer_get_to_zero_terminated_string(PGM_Handle* handle, PGM_Serializer* serializer,
                                                                 PGM_Idx use_compact_list, PGM_Idx indent);
PGM_API void PGM_destroy_serializer(PGM_Serializer* serializer);
PGM_API PGM_ConstDataset* PGM_create_dataset_const_from_writable(PGM_Handle* handle, PGM_WritableDataset const* writable_dataset);
PGM_API PGM_DatasetInfo const* PGM_dataset_writable_get_info(PGM_Handle* handle, PGM_WritableDataset const* dataset);
PGM_API void PGM_dataset_writable_set_buffer(PGM_Handle* handle, PGM_WritableDataset* dataset, char const* component,
                                             PGM_Idx* indptr, void* data);
PGM_API void PGM_dataset_writable_set_attribute_buffer(PGM_Handle* handle, PGM_WritableDataset* dataset, char const* component,
                                                       char const* attribute, void* data);

#ifdef __cplusplus
}
#endif
#endif /* PGM_B200_CAPI_H */
