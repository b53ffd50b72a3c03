// nn.cuh -- device ResNet agent (nn.cu) as seen by api.cu.
#pragma once
#include "handle.cuh"

bool nn_ready(const tz_handle* h);
void nn_free(tz_handle* h);
// policy/value/uncertainty of the queued leaf positions -> d.logits / d.value / d.variance
int nn_forward_queue(tz_handle* h);
