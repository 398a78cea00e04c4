// Host mirror of src/workload.rs:1-69 (grid / block descriptors, ceil, compute_dim, WebGPU limit constants).
#include <sstream>

#include "../../../include/wgpu_mm.hpp"

namespace wgpu_mm {

std::pair<uint32_t, uint32_t> Workload::compute_dim(size_t work_items, WorkloadDim dim) {
    const size_t max_workgroup_size = dim == WorkloadDim::X   ? MAX_WORKGROUP_SIZE_X
                                      : dim == WorkloadDim::Y ? MAX_WORKGROUP_SIZE_Y
                                                              : MAX_WORKGROUP_SIZE_Z;
    const size_t max_workgroup_count = MAX_COMPUTE_WORKGROUPS_PER_DIMENSION;
    if (work_items > max_workgroup_count) {
        const size_t workgroup_size = ceil(work_items, max_workgroup_count);
        const size_t workgroup_count = ceil(work_items, workgroup_size);
        if (workgroup_count > max_workgroup_count || workgroup_size > max_workgroup_size)
            throw Panic("Compute limits exceeded");  // src/workload.rs:60
        return {(uint32_t)workgroup_count, (uint32_t)workgroup_size};
    }
    // one workgroup per work item (src/workload.rs:63-66)
    return {(uint32_t)work_items, 1u};
}

std::string Workload::debug() const {
    std::ostringstream o;
    o << "Workload { count: WorkgroupCount(" << count_.x << ", " << count_.y << ", " << count_.z
      << "), size: WorkgroupSize(" << size_.x << ", " << size_.y << ", " << size_.z << ") }";
    return o.str();
}

std::string KernelSpec::describe() const {
    std::ostringstream o;
    o << "kernel " << name << " (id " << kernel_id << ") workgroup_size=(" << params.workgroup_size[0] << ","
      << params.workgroup_size[1] << "," << params.workgroup_size[2] << ")";
    if (params.absmax != 0.f) o << " absmax=" << params.absmax;
    if (params.batch) o << " batch=" << params.batch;
    if (params.flags) o << " flags=0x" << std::hex << params.flags << std::dec;
    return o.str();
}

}  // namespace wgpu_mm
