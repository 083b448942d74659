//! Golden-vector dumper for the resampler of the reference (closes "sample-level parity unpinned").
//!
//! Untested source: there is no Rust toolchain in the build image.  On a machine that has one, copy this file to
//! `examples/dump_resample_golden.rs` of tphakala/birda and run
//!     cargo run --release --example dump_resample_golden -- /tmp/birda_golden
//! It pushes deterministic inputs through the reference's own `birda::audio::resample` (src/audio/resample.rs:10-92,
//! rubato 4.0.0 `Fft::<f32>::new(from, to, 1024, 1, FixedSync::Both)`) and writes, per case, two raw little-endian
//! f32 files (`<case>.in.f32`, `<case>.out.f32`) plus `cases.txt` (name from_rate to_rate n_in n_out).
//! `python tools/check_rust_golden.py /tmp/birda_golden` then holds this repo's oracle (and, on a GPU box, the CUDA
//! path) to those vectors within the north-star tolerance of 1e-5 relative.
use std::fs::File;
use std::io::{BufWriter, Write};
use std::path::Path;

/// Deterministic test signal: three tones plus LCG noise at about -30 dBFS, identical to `lcg_signal` in
/// tools/check_rust_golden.py (integer LCG, so both sides produce bit-identical f32 input).
fn lcg_signal(n: usize, rate: u32, seed: u64) -> Vec<f32> {
    let mut state = seed;
    let mut out = Vec::with_capacity(n);
    for i in 0..n {
        state = state
            .wrapping_mul(6364136223846793005)
            .wrapping_add(1442695040888963407);
        let noise = ((state >> 40) as f64 / (1u64 << 24) as f64 - 0.5) * 0.0632;
        let t = i as f64 / f64::from(rate);
        let s = 0.08 * (2.0 * std::f64::consts::PI * 1000.0 * t).sin()
            + 0.08 * (2.0 * std::f64::consts::PI * 3217.0 * t).sin()
            + 0.08 * (2.0 * std::f64::consts::PI * 7919.0 * t).sin()
            + noise;
        out.push(s as f32);
    }
    out
}

fn write_f32(path: &Path, data: &[f32]) -> std::io::Result<()> {
    let mut w = BufWriter::new(File::create(path)?);
    for v in data {
        w.write_all(&v.to_le_bytes())?;
    }
    w.flush()
}

fn main() -> Result<(), Box<dyn std::error::Error>> {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "birda_golden".to_string());
    std::fs::create_dir_all(&dir)?;
    // (name, from, to, input frames): the source windows of BASELINE configs 2, 3 and 5 (one window each), a short
    // ragged input (tail rule of resample.rs:58-88) and an input of less than two blocks
    let cases: [(&str, u32, u32, usize); 8] = [
        ("c2_window_44100_48000", 44_100, 48_000, 132_300),
        ("c3_window_48000_32000", 48_000, 32_000, 240_000),
        ("c5_window_22050_48000", 22_050, 48_000, 66_150),
        ("c5_window_96000_48000", 96_000, 48_000, 288_000),
        ("c5_window_16000_48000", 16_000, 48_000, 48_000),
        ("c5_window_32000_48000", 32_000, 48_000, 96_000),
        ("ragged_44100_32000", 44_100, 32_000, 10_007),
        ("short_48000_32000", 48_000, 32_000, 2_000),
    ];
    let mut index = BufWriter::new(File::create(Path::new(&dir).join("cases.txt"))?);
    for (k, (name, from, to, n)) in cases.iter().enumerate() {
        let input = lcg_signal(*n, *from, 1000 + k as u64);
        let output = birda::audio::resample(input.clone(), *from, *to)?;
        write_f32(&Path::new(&dir).join(format!("{name}.in.f32")), &input)?;
        write_f32(&Path::new(&dir).join(format!("{name}.out.f32")), &output)?;
        writeln!(index, "{name} {from} {to} {} {}", input.len(), output.len())?;
    }
    index.flush()?;
    Ok(())
}
