// minimap2_b200/csrc/ksw_plan.h -- a prepared K3 kernel launch: all host-side preparation (queue order, sizes, uploads) is done
// before the first kernel starts so that the kernels of one batch run back to back.
#pragma once
#include <functional>
#include <vector>
#include <cstdint>
#include <cstddef>
struct KswPlan {
	size_t pws_bytes = 0, cigws_bytes = 0;                 // traceback / CIGAR workspace this launch needs
	std::function<void(uint8_t *pws, uint32_t *cigws)> go; // enqueue the kernel on the context's stream
};
void mmb_order_by_cells(std::vector<int> &v, const mmb_ksw_job_t *h_jobs);
