// Internal declarations shared by symbolic.cpp and engine.cu.  The host-side
// circuit description (cdn_group_host, cdn_circuit_host) is part of the public
// C ABI: include/cardoon_b200.h.
#pragma once
#include "cardoon_b200.h"

struct cdn_plan_dev;
/* host copy of the device view (pointers are device pointers once bound) */
const cdn_plan_dev* cdn_plan_host_view(cdn_plan* P);
/* uploads the view struct if it changed; returns its device address */
int cdn_plan_sync_device(cdn_plan* P, cudaStream_t st, const cdn_plan_dev** out);
cdn_ctx* cdn_plan_ctx(cdn_plan* P);
/* remembers the thread team of the last launch (reported by cdn_plan_info[12]) */
void cdn_plan_note_team(cdn_plan* P, int team);
