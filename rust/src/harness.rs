//! `test_harness` (upstream src/harness.rs:170-248) over the C ABI: verify one launch against the CPU reference `mm_ref`
//! with the 1e-3 gate, then 8 warm-up and 10 timed launches with the upstream buffer rotation, print nanoseconds and GFLOPS.
//! The five wgpu touch-points are `device::Device::{new, create_buffer_init, create_compute_pipeline, mm}` and
//! `Buffer::to_cpu`; everything numerical -- the reference product, the error metric, the data distribution -- stays here.
use crate::device::{Buffer, Device, Pipeline};
use crate::gemv::ABSMAX;
use crate::quant::{sint8_dequantize, sint8_quantize};
use crate::{WorkgroupCount, Workload};
use std::time::Instant;

/// The reference's own oracle (src/harness.rs:17-28): fp32, k-sequential, separate multiply and add per step.
fn mm_ref(A: &[f32], B: &[f32], C: &mut [f32], dims: (usize, usize, usize)) {
    let (M, N, K) = dims;
    for (m, c_row) in C.chunks_exact_mut(N).enumerate().take(M) {
        let a_row = &A[m * K..(m + 1) * K];
        for (n, c) in c_row.iter_mut().enumerate() {
            let mut acc = 0f32;
            for (k, a) in a_row.iter().enumerate() {
                acc += a * B[k * N + n];
            }
            *c = acc;
        }
    }
}

/// splitmix64 counter generator: element i of stream `seed` is U[-10, 10) / 50 (src/harness.rs:103-121 draws the same
/// distribution from an unseeded thread_rng).  Bit-identical to the device generator b200mm_buffer_fill_weights and to
/// the C++ mirror, so a failing seed can be replayed on any front end.
fn generate_weight_data(seed: u64, M: usize, N: usize) -> Vec<f32> {
    let mix = |mut x: u64| {
        x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
        x = (x ^ (x >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        x = (x ^ (x >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        x ^ (x >> 31)
    };
    (0..(M * N) as u64)
        .map(|i| {
            let u24 = (mix(seed.wrapping_mul(0xD134_2543_DE82_EF95).wrapping_add(i)) >> 40) as u32;
            let unit = u24 as f32 * (1.0 / 16_777_216.0);
            (unit * 20.0 - 10.0) / 50.0
        })
        .collect()
}

fn harness_seed() -> u64 {
    std::env::var("WGPU_MM_SEED").ok().and_then(|s| s.parse().ok()).unwrap_or(0x5EED)
}

fn rand_gpu_buffer<'d>(device: &'d Device, seed: u64, dims: (usize, usize), return_cpu: bool) -> (Buffer<'d>, Option<Vec<f32>>) {
    let data = generate_weight_data(seed, dims.0, dims.1);
    let buffer = device.create_buffer_init(&data);
    (buffer, if return_cpu { Some(data) } else { None })
}

/// src/harness.rs:123-146: the true absmax is dropped on purpose -- both sides dequantise with gemv::ABSMAX
fn rand_quantized_gpu_buffer<'d>(device: &'d Device, seed: u64, dims: (usize, usize), return_cpu: bool) -> (Buffer<'d>, Option<Vec<u32>>) {
    let data = generate_weight_data(seed, dims.0, dims.1);
    let (quantized, _absmax) = sint8_quantize(&data, dims.0, dims.1);
    let buffer = device.create_buffer_init(&quantized);
    (buffer, if return_cpu { Some(quantized) } else { None })
}

fn is_quantised_kernel(shader: &str) -> bool {
    matches!(shader, "qgemv_1" | "qgemv_sint8")
}

/// src/harness.rs:30-85
fn check(device: &Device, pipeline: &Pipeline, workgroup_count: &WorkgroupCount, dims: (usize, usize, usize), quantized: bool, seed: u64) {
    let (M, N, K) = dims;
    let (A, A_cpu) = rand_gpu_buffer(device, seed + 1, (M, K), true);
    let (B, B_cpu) = if quantized {
        let (B, words) = rand_quantized_gpu_buffer(device, seed + 2, (K, N), true);
        (B, sint8_dequantize(&words.unwrap(), ABSMAX, K, N))
    } else {
        let (B, data) = rand_gpu_buffer(device, seed + 2, (K, N), true);
        (B, data.unwrap())
    };
    // C starts as noise: the kernel must overwrite it (alpha = 1, beta = 0)
    let (C, C_cpu) = rand_gpu_buffer(device, seed + 3, (M, N), true);
    let mut C_cpu = C_cpu.unwrap();
    mm_ref(&A_cpu.unwrap(), &B_cpu, &mut C_cpu, dims);

    device.mm(pipeline, &A, &B, &C, workgroup_count);
    let gpu_out = C.to_cpu();

    let mae = gpu_out.iter().zip(C_cpu.iter()).map(|(g, c)| (g - c).abs()).fold(0f32, f32::max);
    let edge = 16.min(M * N);
    println!("GPU\n{:?}\n...\n{:?}", &gpu_out[..edge], &gpu_out[M * N - edge..]);
    println!("CPU\n{:?}\n...\n{:?}", &C_cpu[..edge], &C_cpu[M * N - edge..]);
    println!("Max Absolute Error: {}", mae);
    if !(mae <= 1e-3) {
        panic!("MAE too high");
    }
}

pub async fn test_harness(workload: Workload, shader: String, dims: (usize, usize, usize), quantize_b: bool) {
    let device = Device::new();
    let (M, N, K) = dims;
    if quantize_b != is_quantised_kernel(&shader) {
        // wgpu's bind-group validation caught a u32 buffer bound where the shader declares f32 (and vice versa)
        panic!("binding 1 type mismatch: quantize_b does not match the kernel's B operand");
    }
    println!("shader: {} on {}", shader, device.name());
    let pipeline = device.create_compute_pipeline(&shader, dims, &workload, if quantize_b { ABSMAX } else { 0.0 });
    let seed = harness_seed();

    check(&device, &pipeline, workload.count(), dims, quantize_b, seed);

    let (A, _) = rand_gpu_buffer(&device, seed + 11, (M, K), false);
    let B = if quantize_b {
        rand_quantized_gpu_buffer(&device, seed + 12, (K, N), false).0
    } else {
        rand_gpu_buffer(&device, seed + 12, (K, N), false).0
    };
    let (C, _) = rand_gpu_buffer(&device, seed + 13, (M, N), false);

    // The upstream loop rotates the roles of the three buffers (src/harness.rs:212-237).  That is only shape-legal when
    // M == N == K and B is f32; otherwise every launch is (A, B, C).
    let rotate = M == N && N == K && !quantize_b;
    let order: [(u8, u8, u8); 10] = [(0, 1, 2), (2, 1, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (2, 1, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (1, 0, 2)];
    let bufs = [&A, &B, &C];
    let submit = |launches: &[(u8, u8, u8)]| {
        for &(a, b, c) in launches {
            let (a, b, c) = if rotate { (a, b, c) } else { (0, 1, 2) };
            device.mm(&pipeline, bufs[a as usize], bufs[b as usize], bufs[c as usize], workload.count());
        }
    };

    //warmup
    submit(&order[..8]);
    let _warmup_res = C.to_cpu();

    let start = Instant::now();
    let (_, kernel_ms) = device.timed(|| submit(&order));
    let _result = C.to_cpu();
    let elapsed = start.elapsed();

    let nanos = elapsed.as_nanos();
    println!("{} ns", nanos);
    let flops = M * N * K * 2 * 10;
    let gflops = (flops as f64 / 1e9) / (nanos as f64 / 1e9);
    println!("{} GFLOPS", gflops);
    println!("{} GFLOPS (CUDA events around the 10 launches, no read-back)", flops as f64 / 1e9 / (kernel_ms as f64 / 1e3));
}
