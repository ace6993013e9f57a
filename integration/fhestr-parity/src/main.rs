//! fhestr-parity: golden vectors out of tfhe-rs 0.5.2 for the engine's kernels, and key export.
//!
//! AUTHORED, NOT COMPILED (no Rust toolchain in the build image).  API names follow tfhe-rs 0.5.2's
//! `core_crypto` prelude as the author recalls them; a maintainer with cargo may need to adjust an import.
//!
//! Sub-commands
//!   fixture <out.bin>                  keys + (input LWE, keyswitched LWE, PBS output) triples + LUT accumulators
//!   export-keys <server_key.bincode> <out.bin>
//!                                      a bincode `tfhe::integer::ServerKey` (what the reference holds,
//!                                      /root/reference/src/server_key/mod.rs:13-16, serialisable through the serde
//!                                      derive of /root/reference/src/client_key.rs:9) -> the same flat file with
//!                                      bsk_std and ksk only (no secrets, no triples)
//!
//! Flat file "FHESTRFX" v1, little endian (reader: fhestring_b200/fixtures.py):
//!   magic[8] = "FHESTRFX", u32 version = 1,
//!   i32 n, N, k, pbs_base_log, pbs_level, ks_base_log, ks_level, delta_log,
//!   u32 has_secrets, u32 n_luts, u32 n_triples,
//!   if has_secrets: s_lwe[n] u8, s_glwe[k*N] u8
//!   bsk_std[n][pbs_level][k+1][k+1][N] u64        standard domain, what fhestr_load_keys takes
//!   ksk[k*N][ks_level][n+1] u64                   level 1 first
//!   n_luts   x { table[16] u8, body[N] u64 }      accumulator body polynomial of generate_lookup_table
//!   n_triples x { u32 lut, in[k*N+1] u64, ks[n+1] u64, out[k*N+1] u64 }
//!
//! What each section pins (tests/test_tfhe_rs_fixture.py):
//!   ksk + in -> ks      K0/K1 keyswitch, bit-exact (signed decomposer, level order, sign convention: SURVEY App. A.4/A.5)
//!   table -> body       K5 LUT generation, bit-exact (box size, half-box rotation, negated first half-box: SURVEY 2.5)
//!   ks, body -> out     K2 mod-switch + K3 blind rotation + K4 sample extract: decrypts identically, phase within the
//!                       stated torus tolerance of tfhe-rs' own f64 FFT PBS (different FFT rounding, same algorithm)
//!   bsk_std             the key hand-over itself: the engine's Fourier key is derived from these words
use std::fs::File;
use std::io::{BufWriter, Write};

use tfhe::core_crypto::prelude::*;
use tfhe::shortint::parameters::PARAM_MESSAGE_2_CARRY_2_KS_PBS;

fn w32(f: &mut impl Write, v: u32) { f.write_all(&v.to_le_bytes()).unwrap(); }
fn wi32(f: &mut impl Write, v: i32) { f.write_all(&v.to_le_bytes()).unwrap(); }
fn w64s(f: &mut impl Write, v: &[u64]) { for x in v { f.write_all(&x.to_le_bytes()).unwrap(); } }

struct Dims { n: usize, big_n: usize, k: usize, pbs_base_log: usize, pbs_level: usize, ks_base_log: usize, ks_level: usize }

fn header(f: &mut impl Write, d: &Dims, has_secrets: u32, n_luts: u32, n_triples: u32) {
    f.write_all(b"FHESTRFX").unwrap();
    w32(f, 1);
    for v in [d.n, d.big_n, d.k, d.pbs_base_log, d.pbs_level, d.ks_base_log, d.ks_level, 59usize] { wi32(f, v as i32); }
    w32(f, has_secrets); w32(f, n_luts); w32(f, n_triples);
}

/// the 16-entry function tables the fixture covers: identity, the bivariate-eq table of fheasciichar.rs:36, x -> 3x+1,
/// the carry / message extraction tables of the radix ops, the nz table the comparison uses on the padding-bit half
fn tables() -> Vec<[u8; 16]> {
    let mut t = Vec::new();
    t.push(core::array::from_fn(|x| x as u8));
    t.push(core::array::from_fn(|x| ((x >> 2) == (x & 3)) as u8));
    t.push(core::array::from_fn(|x| ((3 * x + 1) % 16) as u8));
    t.push(core::array::from_fn(|x| (x >> 2) as u8));
    t.push(core::array::from_fn(|x| (x & 3) as u8));
    t.push(core::array::from_fn(|x| (x != 0) as u8));
    t
}

/// shortint's accumulator for f over the 16 block values (message 4 x carry 4), built with shortint's OWN
/// generate_lookup_table so that K5 is pinned against the code path the reference actually runs
fn accumulator_body(sks: &tfhe::shortint::ServerKey, table: &[u8; 16]) -> Vec<u64> {
    let tb = *table;
    let lut = sks.generate_lookup_table(move |x| tb[(x % 16) as usize] as u64);
    // LookupTableOwned { acc: GlweCiphertextOwned<u64>, degree }: trivial GLWE, the body is the last polynomial
    lut.acc.get_body().as_ref().to_vec()
}

fn fixture(out: &str) {
    let p = PARAM_MESSAGE_2_CARRY_2_KS_PBS;
    let d = Dims { n: p.lwe_dimension.0, big_n: p.polynomial_size.0, k: p.glwe_dimension.0, pbs_base_log: p.pbs_base_log.0,
                   pbs_level: p.pbs_level.0, ks_base_log: p.ks_base_log.0, ks_level: p.ks_level.0 };
    // keys through the SAME entry point the reference uses (client_key.rs:31: gen_keys_radix -> shortint keys); the
    // standard-domain BSK is recovered from the Fourier one with tfhe-rs' own inverse transform (see bsk_to_standard)
    let (cks, sks) = tfhe::shortint::gen_keys(p);
    let (small, glwe) = (cks.small_lwe_secret_key(), cks.glwe_secret_key.clone());   // 0.5.2 field / accessor names
    let big = glwe.clone().into_lwe_secret_key();

    let tabs = tables();
    let n_triples = 64u32;
    let mut f = BufWriter::new(File::create(out).unwrap());
    header(&mut f, &d, 1, tabs.len() as u32, n_triples);
    f.write_all(&small.as_ref().iter().map(|&b| b as u8).collect::<Vec<_>>()).unwrap();
    f.write_all(&big.as_ref().iter().map(|&b| b as u8).collect::<Vec<_>>()).unwrap();
    w64s(&mut f, &bsk_to_standard(&sks, &d));
    w64s(&mut f, sks.key_switching_key.as_ref());
    let mut bodies = Vec::new();
    for t in &tabs {
        let body = accumulator_body(&sks, t);
        f.write_all(t).unwrap();
        w64s(&mut f, &body);
        bodies.push(body);
    }
    // triples: encrypt a block value (values 0..31: the padding-bit half included through a raw plaintext), keyswitch
    // and bootstrap with tfhe-rs' own core_crypto primitives on the keys above
    let fourier_bsk = match &sks.bootstrapping_key {
        tfhe::shortint::server_key::ShortintBootstrappingKey::Classic(b) => b,
        _ => panic!("classic PBS expected"),
    };
    let mut enc = tfhe::core_crypto::commons::generators::EncryptionRandomGenerator::<ActivatedRandomGenerator>::new(
        tfhe::core_crypto::seeders::new_seeder().seed(), tfhe::core_crypto::seeders::new_seeder().as_mut());
    for i in 0..n_triples as usize {
        let lut = i % tabs.len();
        let value = (i as u64 * 7 + 3) % 32;
        let mut input = LweCiphertext::new(0u64, big.lwe_dimension().to_lwe_size(), CiphertextModulus::new_native());
        encrypt_lwe_ciphertext(&big, &mut input, Plaintext(value << 59), p.glwe_modular_std_dev, &mut enc);
        let mut ks = LweCiphertext::new(0u64, small.lwe_dimension().to_lwe_size(), CiphertextModulus::new_native());
        keyswitch_lwe_ciphertext(&sks.key_switching_key, &input, &mut ks);
        let mut acc = GlweCiphertext::new(0u64, p.glwe_dimension.to_glwe_size(), p.polynomial_size, CiphertextModulus::new_native());
        acc.get_mut_body().as_mut().copy_from_slice(&bodies[lut]);
        let mut outp = LweCiphertext::new(0u64, big.lwe_dimension().to_lwe_size(), CiphertextModulus::new_native());
        programmable_bootstrap_lwe_ciphertext(&ks, &mut outp, &acc, fourier_bsk);
        w32(&mut f, lut as u32);
        w64s(&mut f, input.as_ref());
        w64s(&mut f, ks.as_ref());
        w64s(&mut f, outp.as_ref());
    }
    f.flush().unwrap();
    println!("wrote {out}: n={} N={} {} luts {} triples", d.n, d.big_n, tabs.len(), n_triples);
}

/// Fourier BSK -> standard domain with tfhe-rs' own FFT plan, polynomial by polynomial:
/// [n][pbs_level][k+1 rows][k+1 cols][N].  The Fourier key is all tfhe-rs keeps (shortint::ServerKey.bootstrapping_key);
/// its layout is concrete-fft's, which is why the engine never reads it in place.  The
/// round trip loses nothing the CPU path used: the CPU multiplies by exactly these Fourier values.
fn bsk_to_standard(sks: &tfhe::shortint::ServerKey, d: &Dims) -> Vec<u64> {
    use tfhe::core_crypto::fft_impl::fft64::math::fft::Fft;
    let fbsk = match &sks.bootstrapping_key {
        tfhe::shortint::server_key::ShortintBootstrappingKey::Classic(b) => b,
        _ => panic!("classic PBS expected"),
    };
    let fft = Fft::new(PolynomialSize(d.big_n));
    let fft = fft.as_view();
    let mut mem = dyn_stack::GlobalPodBuffer::new(fft.backward_scratch().unwrap());
    let mut out = vec![0u64; d.n * d.pbs_level * (d.k + 1) * (d.k + 1) * d.big_n];
    let mut o = 0;
    // FourierLweBootstrapKey::as_view().into_ggsw_iter(): one Fourier GGSW per input key bit, in key order;
    // each GGSW: levels (tfhe-rs stores the LAST level first in memory: reverse here so that level 1 comes first,
    // which is what fhestr_load_keys takes), rows, columns, polynomials of N/2 c64
    for ggsw in fbsk.as_view().into_ggsw_iter() {
        let mut levels: Vec<_> = ggsw.into_levels().collect();
        levels.sort_by_key(|l| l.decomposition_level().0);
        for level in levels {
            for row in level.into_rows() {
                for poly in row.data().chunks_exact(d.big_n / 2) {
                    let mut std_poly = Polynomial::new(0u64, PolynomialSize(d.big_n));
                    fft.backward_as_torus(std_poly.as_mut_view(), FourierPolynomial { data: poly },
                                          dyn_stack::PodStack::new(&mut mem));
                    out[o..o + d.big_n].copy_from_slice(std_poly.as_ref());
                    o += d.big_n;
                }
            }
        }
    }
    assert_eq!(o, out.len());
    out
}

fn export_keys(input: &str, out: &str) {
    let bytes = std::fs::read(input).unwrap();
    let isk: tfhe::integer::ServerKey = bincode::deserialize(&bytes).unwrap();
    let sks: tfhe::shortint::ServerKey = isk.into();          // integer::ServerKey is a newtype over the shortint key
    let p = PARAM_MESSAGE_2_CARRY_2_KS_PBS;
    let d = Dims { n: sks.key_switching_key.output_key_lwe_dimension().0, big_n: p.polynomial_size.0, k: p.glwe_dimension.0,
                   pbs_base_log: p.pbs_base_log.0, pbs_level: p.pbs_level.0,
                   ks_base_log: sks.key_switching_key.decomposition_base_log().0,
                   ks_level: sks.key_switching_key.decomposition_level_count().0 };
    let mut f = BufWriter::new(File::create(out).unwrap());
    header(&mut f, &d, 0, 0, 0);
    w64s(&mut f, &bsk_to_standard(&sks, &d));
    w64s(&mut f, sks.key_switching_key.as_ref());
    f.flush().unwrap();
    println!("wrote {out}");
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    match a.get(1).map(|s| s.as_str()) {
        Some("fixture") => fixture(&a[2]),
        Some("export-keys") => export_keys(&a[2], &a[3]),
        _ => eprintln!("usage: fhestr-parity fixture <out.bin> | export-keys <server_key.bincode> <out.bin>"),
    }
}
