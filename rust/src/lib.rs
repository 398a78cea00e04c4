//! Crate root with the B200 back-end.  UNCOMPILED in this repository (no Rust toolchain in the build image); kept in step with
//! the C headers by tests/test_host.py::test_rust_shim_binds_every_exported_symbol.
//!
//! Public surface = upstream `src/lib.rs:1-10`: `gemm`, `gemv`, `quant` modules, and `test_harness`, `Workload`,
//! `WorkgroupCount`, `WorkgroupSize`, `WorkloadDim` re-exported at the root.  What changed underneath:
//!   * `wgpu` (device, buffers, pipeline, dispatch, read-back)  ->  `ffi` over libb200mm.so, wrapped by `device`;
//!   * `tera` (WGSL templating)  ->  `tera` below, a two-type stand-in so the entry points keep their signatures
//!     `fn(&mut Tera, &mut Context) -> (Workload, String)`; the `String` is now a kernel name, not WGSL text;
//!   * `rand`  ->  a seeded counter generator (bit-identical to the CUDA-side generator, so 16384^2 operands can also be
//!     produced on the device).
#![allow(non_snake_case)]
pub mod device;
pub mod ffi;
pub mod gemm;
pub mod gemv;
mod harness;
mod launch_shape;
pub mod quant;
pub mod tera;

pub use harness::*;
pub use launch_shape::*;
