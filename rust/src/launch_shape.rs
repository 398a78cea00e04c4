//! `Workload`, `WorkgroupCount`, `WorkgroupSize`, `WorkloadDim` (upstream src/workload.rs:1-69).  The types are plain data and
//! stay in Rust; the limit arithmetic is shared with the C++ mirror through the C entry points so both sides cannot drift.
use crate::ffi;

/// gridDim
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct WorkgroupCount(pub u32, pub u32, pub u32);
/// blockDim
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct WorkgroupSize(pub u32, pub u32, pub u32);

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct Workload {
    count: WorkgroupCount,
    size: WorkgroupSize,
}

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum WorkloadDim {
    X,
    Y,
    Z,
}

impl Workload {
    pub const MAX_WORKGROUP_SIZE_X: usize = 256;
    pub const MAX_WORKGROUP_SIZE_Y: usize = 256;
    pub const MAX_WORKGROUP_SIZE_Z: usize = 64;
    pub const MAX_COMPUTE_WORKGROUPS_PER_DIMENSION: usize = 65535;

    pub fn new(count: WorkgroupCount, size: WorkgroupSize) -> Self {
        Workload { count, size }
    }
    pub fn count(&self) -> &WorkgroupCount {
        &self.count
    }
    pub fn size(&self) -> &WorkgroupSize {
        &self.size
    }
    pub fn ceil(num: usize, div: usize) -> usize {
        unsafe { ffi::wgpumm_workload_ceil(num, div) }
    }
    /// (workgroup_count, workgroup_size) for one dimension; panics with "Compute limits exceeded" (src/workload.rs:60)
    pub fn compute_dim(work_items: usize, dim: WorkloadDim) -> (u32, u32) {
        let (mut count, mut size) = (0u32, 0u32);
        let axis = match dim {
            WorkloadDim::X => 0,
            WorkloadDim::Y => 1,
            WorkloadDim::Z => 2,
        };
        let rc = unsafe { ffi::wgpumm_compute_dim(work_items, axis, &mut count, &mut size) };
        if rc != 0 {
            panic!("Compute limits exceeded");
        }
        (count, size)
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn compute_dim_small_counts_map_to_one_item_per_workgroup() {
        assert_eq!(Workload::compute_dim(1000, WorkloadDim::X), (1000, 1));
        assert_eq!(Workload::compute_dim(65535, WorkloadDim::Y), (65535, 1));
    }

    #[test]
    fn compute_dim_folds_large_counts_into_the_workgroup_size() {
        assert_eq!(Workload::compute_dim(65536, WorkloadDim::X), (32768, 2));
        assert_eq!(Workload::compute_dim(1 << 20, WorkloadDim::X), (61681, 17));
    }

    #[test]
    #[should_panic(expected = "Compute limits exceeded")]
    fn compute_dim_panics_past_the_limits() {
        Workload::compute_dim(65535 * 64 + 1, WorkloadDim::Z);
    }

    #[test]
    fn ceil_rounds_up() {
        assert_eq!(Workload::ceil(1024, 16), 64);
        assert_eq!(Workload::ceil(1025, 16), 65);
    }
}
