//! Known-answer generator: the reference's own `sequence -> jitter -> synthesize` chain (src/lib.rs:936-953, 781-801,
//! 582-600) on the phoneme lists of SURVEY.md Appendix B, written as raw little-endian f32 plus a word-wise FNV-1a.
//! UNCOMPILED HERE (no rustc in the build image).  Output: <dir>/<name>.f32 and <dir>/index.txt (name n fnv).
use grail_rs::*;
use std::io::Write;

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| ".".into());
    let v = voices::generic();
    let (s, a, e) = (Phoneme::Silence, Phoneme::A, Phoneme::E);
    let cases: Vec<(&str, Vec<Phoneme>, u32)> = vec![
        ("sil_a", vec![s, a], 0),
        ("ten", vec![s, e, a, a, e, a, a, e, a, a], 0),
        ("ten_seed12345", vec![s, e, a, a, e, a, a, e, a, a], 12345),
        ("sil_sil_sil_a", vec![s, s, s, a], 0),
        ("e_a_seed12345", vec![e, a], 12345),
        ("stop_glide_a_sil_e", vec![Phoneme::Stop, Phoneme::Glide, a, s, e], 7),
    ];
    let mut index = std::fs::File::create(format!("{dir}/index.txt")).unwrap();
    for (name, ph, seed) in cases {
        // Intonator (src/lib.rs:1057-1075) as it stands: length 0.5, blend 0.5, the voice's centre frequency
        let audio: Vec<f32> = ph
            .into_iter()
            .map(|p| PhonemeElem { phoneme: p, length: 0.5, blend_length: 0.5, frequency: v.center_frequency })
            .select(v)
            .sequence(v)
            .jitter(seed, v)
            .synthesize()
            .collect();
        let mut h: u32 = 2166136261;
        let mut raw = Vec::with_capacity(audio.len() * 4);
        for x in &audio {
            h = (h ^ x.to_bits()).wrapping_mul(16777619);
            raw.extend_from_slice(&x.to_le_bytes());
        }
        std::fs::write(format!("{dir}/{name}.f32"), raw).unwrap();
        writeln!(index, "{name} {} {h:08x}", audio.len()).unwrap();
    }
}
