// trainer.cu -- whole-step executor (placeholder: entry points exist, executor lands next).
#include "common.cuh"
#define NOTYET(name) { cenn_set_error(name ": fused executor not available in this build"); return 1; }
extern "C" {
int cenn_trainer_create(cenn_state *, const cenn_trainer_config *, cenn_trainer **) NOTYET("cenn_trainer_create")
int cenn_trainer_destroy(cenn_trainer *) { return 0; }
int cenn_trainer_param_count(cenn_trainer *, int, int64_t *) NOTYET("cenn_trainer_param_count")
int cenn_trainer_set_params_host(cenn_trainer *, int, const float *) NOTYET("cenn_trainer_set_params_host")
int cenn_trainer_get_params_host(cenn_trainer *, int, float *) NOTYET("cenn_trainer_get_params_host")
int cenn_trainer_get_grads_host(cenn_trainer *, int, float *) NOTYET("cenn_trainer_get_grads_host")
int cenn_trainer_bn_stat_count(cenn_trainer *, int, int64_t *) NOTYET("cenn_trainer_bn_stat_count")
int cenn_trainer_set_bn_stats_host(cenn_trainer *, int, const float *) NOTYET("cenn_trainer_set_bn_stats_host")
int cenn_trainer_get_bn_stats_host(cenn_trainer *, int, float *) NOTYET("cenn_trainer_get_bn_stats_host")
int cenn_trainer_step_host(cenn_trainer *, const float *, const float *, const uint8_t *, float *) NOTYET("cenn_trainer_step_host")
int cenn_trainer_step_device(cenn_trainer *, const float *, const float *, const uint8_t *) NOTYET("cenn_trainer_step_device")
int cenn_trainer_read_losses(cenn_trainer *, float *) NOTYET("cenn_trainer_read_losses")
int cenn_trainer_grad_buffer(cenn_trainer *, int, float **, int64_t *) NOTYET("cenn_trainer_grad_buffer")
int cenn_trainer_step_phase(cenn_trainer *, int, const float *, const float *, const uint8_t *) NOTYET("cenn_trainer_step_phase")
int cenn_trainer_generator_forward_host(cenn_trainer *, const float *, float *, int) NOTYET("cenn_trainer_generator_forward_host")
int cenn_trainer_fetch_host(cenn_trainer *, const char *, float *, int64_t, int64_t *) NOTYET("cenn_trainer_fetch_host")
int cenn_trainer_kernel_launches_per_step(cenn_trainer *, int64_t *) NOTYET("cenn_trainer_kernel_launches_per_step")
}
