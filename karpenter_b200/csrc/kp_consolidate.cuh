// kp_consolidate.cuh -- multi-node consolidation search (filled in below kp_api.cu's handle definition)
#pragma once
#include "kp_solve.cuh"
struct kp_handle;
static int kp_consolidate_impl(kp_handle* h, const kp_problem* cluster, const kp_consol_input* in,
                               kp_consol_result* out);
