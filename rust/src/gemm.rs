//! SGEMM entry points with the upstream signatures (src/gemm.rs:9-150): each one fills the context with the constants its
//! kernel is built from, derives the dispatch geometry from M, N, K, and returns `(Workload, String)`.  The `String` used to
//! be rendered WGSL; now it is the kernel name that `test_harness` resolves with `kernel_id_of`.  For the faithful ports
//! (gemm_1 .. gemm3) the Workload is the CUDA grid / block; the B200-native kernels (sgemm_simt, sgemm_tc3x) derive their
//! own launch configuration and the Workload only documents it.
//!
//! One table describes all kernels; the public functions are one-line lookups, so the geometry of a kernel is written
//! down exactly once (and tests/test_host.py compares it with the C++ mirror, csrc/host/entry_points.cc).
use crate::ffi::*;
use crate::tera::{Context, Tera};
use crate::{WorkgroupCount, WorkgroupSize, Workload};

const M: usize = 1024;
const N: usize = 1024;
const K: usize = 1024;

pub fn insert_matrix_dims(context: &mut Context) -> (usize, usize, usize) {
    insert_matrix_dims_with(context, (M, N, K))
}

/// Same with caller-chosen dimensions (the upstream constants are compile-time; BASELINE configs need 4096^3 and 16384^3).
pub fn insert_matrix_dims_with(context: &mut Context, dims: (usize, usize, usize)) -> (usize, usize, usize) {
    context.insert("M", &dims.0);
    context.insert("N", &dims.1);
    context.insert("K", &dims.2);
    dims
}

/// How a kernel's dispatch grid follows from the problem: `x` and `y` are (dimension, divisor) pairs.
#[derive(Clone, Copy)]
enum Dim {
    M,
    N,
    /// M * N (1-D dispatch over output blocks)
    MN,
    One,
}

pub(crate) struct Entry {
    pub name: &'static str,
    pub kernel_id: i32,
    /// tile constants inserted into the context under their upstream names
    consts: &'static [(&'static str, usize)],
    size: (u32, u32, u32),
    /// grid.x = ceil(dim / div), grid.y likewise
    grid_x: (Dim, usize),
    grid_y: (Dim, usize),
}

#[rustfmt::skip]
pub(crate) const ENTRIES: &[Entry] = &[
    // ---- wired by the reference (src/gemm.rs:16-150) ----
    Entry { name: "gemm_1",  kernel_id: B200MM_K_GEMM_1,  consts: &[], size: (16, 16, 1), grid_x: (Dim::M, 16), grid_y: (Dim::N, 16) },
    Entry { name: "gemm_1v", kernel_id: B200MM_K_GEMM_1V, consts: &[], size: (16, 4, 1),  grid_x: (Dim::M, 16), grid_y: (Dim::N, 16) },
    Entry { name: "gemm_2",  kernel_id: B200MM_K_GEMM_2,  consts: &[], size: (256, 1, 1), grid_x: (Dim::M, 16), grid_y: (Dim::N, 16) },
    Entry { name: "gemm_3",  kernel_id: B200MM_K_GEMM_3,  consts: &[("BLOCKSIZE", 16)], size: (256, 1, 1), grid_x: (Dim::M, 16), grid_y: (Dim::N, 16) },
    Entry { name: "gemm_4",  kernel_id: B200MM_K_GEMM_4,  consts: &[("BM", 16), ("BN", 16), ("BK", 8), ("TM", 2)], size: (128, 1, 1),
            grid_x: (Dim::N, 16), grid_y: (Dim::M, 16) },
    Entry { name: "gemm_5",  kernel_id: B200MM_K_GEMM_5,  consts: &[("BM", 32), ("BN", 32), ("BK", 16), ("TM", 4), ("TN", 4)], size: (64, 1, 1),
            grid_x: (Dim::N, 32), grid_y: (Dim::M, 32) },
    // ---- orphan shaders the reference ships without an entry point ----
    Entry { name: "gemm_wonnx", kernel_id: B200MM_K_GEMM_WONNX, consts: &[], size: (256, 1, 1), grid_x: (Dim::MN, 16 * 256), grid_y: (Dim::One, 1) },
    Entry { name: "bram",    kernel_id: B200MM_K_BRAM,    consts: &[], size: (8, 8, 1),   grid_x: (Dim::M, 4 * 8),  grid_y: (Dim::N, 4 * 8) },
    Entry { name: "bram8x8", kernel_id: B200MM_K_BRAM8X8, consts: &[], size: (4, 8, 1),   grid_x: (Dim::M, 4 * 4),  grid_y: (Dim::N, 4 * 8) },
    Entry { name: "gemm3",   kernel_id: B200MM_K_GEMM3,   consts: &[], size: (16, 16, 1), grid_x: (Dim::N, 8 * 16), grid_y: (Dim::M, 4 * 16) },
    // ---- B200-native (north_star): Workload advisory ----
    Entry { name: "sgemm_simt", kernel_id: B200MM_K_SGEMM_SIMT, consts: &[("BM", 128), ("BN", 128), ("BK", 16), ("TM", 8), ("TN", 8)],
            size: (256, 1, 1), grid_x: (Dim::M, 128), grid_y: (Dim::N, 128) },
    Entry { name: "sgemm_tc3x", kernel_id: B200MM_K_SGEMM_TC3X, consts: &[("BM", 128), ("BN", 256), ("BK", 32)],
            size: (256, 1, 1), grid_x: (Dim::One, 1), grid_y: (Dim::One, 1) },
];

pub(crate) fn entry(name: &str) -> &'static Entry {
    ENTRIES
        .iter()
        .chain(crate::gemv::ENTRIES.iter())
        .find(|e| e.name == name)
        .unwrap_or_else(|| panic!("Template '{}' not found", name)) // what tera's render().unwrap() reported
}

/// kernel name (the `String` an entry point returns) -> b200mm_kernel_id
pub fn kernel_id_of(shader: &str) -> i32 {
    entry(shader).kernel_id
}

impl Entry {
    pub(crate) const fn new(name: &'static str, kernel_id: i32, size: (u32, u32, u32), n_per_workgroup: usize) -> Entry {
        Entry { name, kernel_id, consts: &[], size, grid_x: (Dim::N, n_per_workgroup), grid_y: (Dim::One, 1) }
    }

    fn extent(dim: Dim, m: usize, n: usize) -> usize {
        match dim {
            Dim::M => m,
            Dim::N => n,
            Dim::MN => m * n,
            Dim::One => 1,
        }
    }

    /// fills the context and builds the Workload; `context` must already hold M, N, K (insert_matrix_dims)
    pub(crate) fn emit(&self, tera: &mut Tera, context: &mut Context) -> (Workload, String) {
        tera.add_raw_template(self.name, "").unwrap();
        let (m, n, _k) = (context.require("M") as usize, context.require("N") as usize, context.require("K") as usize);
        for (key, value) in self.consts {
            context.insert(key, value);
        }
        let mut count = WorkgroupCount(
            Workload::ceil(Self::extent(self.grid_x.0, m, n), self.grid_x.1) as u32,
            Workload::ceil(Self::extent(self.grid_y.0, m, n), self.grid_y.1) as u32,
            1,
        );
        if self.kernel_id == B200MM_K_SGEMM_TC3X {
            // persistent kernel: one CTA per SM (148 on B200), capped by the number of 128 x 256 tiles
            let tiles = Workload::ceil(m, 128) * Workload::ceil(n, 256);
            count = WorkgroupCount(tiles.min(148) as u32, 1, 1);
        }
        let workload = Workload::new(count, WorkgroupSize(self.size.0, self.size.1, self.size.2));
        context.insert("workgroup_size_x", &workload.size().0);
        context.insert("workgroup_size_y", &workload.size().1);
        context.insert("workgroup_size_z", &workload.size().2);
        log::debug!("workload: {:?}", workload);
        (workload, self.name.to_string())
    }
}

macro_rules! entry_point {
    ($($name:ident),*) => {$(
        pub fn $name(tera: &mut Tera, context: &mut Context) -> (Workload, String) {
            entry(stringify!($name)).emit(tera, context)
        }
    )*};
}
entry_point!(gemm_1, gemm_1v, gemm_2, gemm_3, gemm_4, gemm_5, gemm_wonnx, bram, bram8x8, gemm3, sgemm_simt, sgemm_tc3x);

#[cfg(test)]
mod tests {
    use crate::test_harness;

    use super::*;

    macro_rules! gemm_test {
        ($test_name:ident, $gemm_function:ident) => {
            #[tokio::test]
            pub async fn $test_name() {
                let _ = env_logger::builder().is_test(true).try_init();
                let mut tera = Tera::default();
                let mut context = Context::new();
                let dims = insert_matrix_dims(&mut context);
                let (workload, shader) = $gemm_function(&mut tera, &mut context);
                test_harness(workload, shader, dims, false).await;
            }
        };
    }

    // the upstream list (src/gemm.rs:172-177)
    gemm_test!(test_gemm_1, gemm_1);
    gemm_test!(test_gemm_1v, gemm_1v);
    gemm_test!(test_gemm_2, gemm_2);
    gemm_test!(test_gemm_3, gemm_3);
    gemm_test!(test_gemm_4, gemm_4);
    gemm_test!(test_gemm_5, gemm_5);
    // orphan shaders + the B200-native kernels
    gemm_test!(test_gemm_wonnx, gemm_wonnx);
    gemm_test!(test_bram, bram);
    gemm_test!(test_bram8x8, bram8x8);
    gemm_test!(test_gemm3, gemm3);
    gemm_test!(test_sgemm_simt, sgemm_simt);
    gemm_test!(test_sgemm_tc3x, sgemm_tc3x);

    #[test]
    fn geometry_at_the_reference_shape() {
        let mut tera = Tera::default();
        let mut context = Context::new();
        insert_matrix_dims(&mut context);
        let (w, s) = gemm_5(&mut tera, &mut context);
        assert_eq!((s.as_str(), *w.count(), *w.size()), ("gemm_5", WorkgroupCount(32, 32, 1), WorkgroupSize(64, 1, 1)));
        assert_eq!(context.int("TM"), Some(4));
        let (w, _) = gemm_1v(&mut tera, &mut context);
        assert_eq!((*w.count(), *w.size()), (WorkgroupCount(64, 64, 1), WorkgroupSize(16, 4, 1)));
        let (w, _) = gemm_wonnx(&mut tera, &mut context);
        assert_eq!((*w.count(), *w.size()), (WorkgroupCount(256, 1, 1), WorkgroupSize(256, 1, 1)));
        let (w, _) = sgemm_tc3x(&mut tera, &mut context);
        assert_eq!(*w.count(), WorkgroupCount(32, 1, 1));
    }

    #[tokio::test]
    #[should_panic(expected = "MAE too high")]
    pub async fn test_gate_panics_on_a_short_dispatch() {
        // an 8 x 8 corner of C only: the rest keeps its noise and the 1e-3 gate must trip (src/harness.rs:82-84)
        let mut tera = Tera::default();
        let mut context = Context::new();
        let dims = insert_matrix_dims_with(&mut context, (64, 64, 64));
        let (_, shader) = gemm_1(&mut tera, &mut context);
        let short = Workload::new(WorkgroupCount(1, 1, 1), WorkgroupSize(8, 8, 1));
        test_harness(short, shader, dims, false).await;
    }
}
