//! UNBUILT IN THIS ENVIRONMENT (no Rust toolchain).
//! The draw contract of DESIGN.md §2 as plain functions, to replace the `rng.random_*` calls inside the
//! reference's example systems so that a Bevy run consumes the same stream as the CUDA engine.

const M0: u32 = 0xD251_1F53;
const M1: u32 = 0xCD9E_8D57;
const W0: u32 = 0x9E37_79B9;
const W1: u32 = 0xBB67_AE85;

pub fn philox4x32_10(mut c: [u32; 4], mut k: [u32; 2]) -> [u32; 4] {
    for _ in 0..10 {
        let p0 = (M0 as u64) * (c[0] as u64);
        let p1 = (M1 as u64) * (c[2] as u64);
        c = [((p1 >> 32) as u32) ^ c[1] ^ k[0], p1 as u32, ((p0 >> 32) as u32) ^ c[3] ^ k[1], p0 as u32];
        k = [k[0].wrapping_add(W0), k[1].wrapping_add(W1)];
    }
    c
}

#[derive(Clone, Copy)]
pub struct Stream { k: [u32; 2] }

impl Stream {
    pub fn new(seed: u64, replica: u32, tag: &[u8; 4]) -> Self {
        let tag = u32::from_be_bytes(*tag);
        Self { k: [(seed as u32) ^ replica, ((seed >> 32) as u32) ^ tag] }
    }
    pub fn block(&self, c0: u32, c1: u32, step: u32, c3: u32) -> [u32; 4] { philox4x32_10([c0, c1, step, c3], self.k) }
}

pub fn threshold(p: f64) -> u64 {
    if !(p > 0.0) { 0 } else if p >= 1.0 { 1u64 << 32 } else { (p * 4294967296.0) as u64 }
}
pub fn bern(u: u32, thr: u64) -> bool { (u as u64) < thr }
pub fn below(u: u32, n: u32) -> u32 { (((u as u64) * (n as u64)) >> 32) as u32 }

/// examples/coin_toss.rs:70  `rng.random_bool(ODDS_HEADS)`
pub fn coin_toss(seed: u64, replica: u32, uid: u32, step: u32, odds: f64) -> bool {
    let b = Stream::new(seed, replica, b"COIN").block(uid >> 2, 0, step, 0);
    bern(b[(uid & 3) as usize], threshold(odds))
}

/// examples/forest_fire.rs:309  spread draw for healthy target `t` and burning source `s`;
/// `dir` = direction of s as seen from t (N(0,-1)=0, W=1, E=2, S=3); `base_cells` = W*H.
pub fn forest_spread(seed: u64, replica: u32, t: u32, s: u32, dir: u32, base_cells: u32, step: u32, p: f64) -> bool {
    let st = Stream::new(seed, replica, b"FSPR");
    if s < base_cells { bern(st.block(t, 0, step, 0)[dir as usize], threshold(p)) } else { bern(st.block(t, s, step, 1)[0], threshold(p)) }
}

/// Block lottery, k = 8: floor(P(Binomial(256, 1/256) <= j) * 2^64), j = 0..20
/// (the tables for k = 5..7 are in tools/gen_lottery_tables.py; the shim only needs this one).
pub const LOTTERY_CDF_K8: [u64; 21] = [
    0x5dfe2e83aaa155df, 0xbc5ab99263fc066e, 0xeb88ff19c0a95eb6,
    0xfb334c65d160bff4, 0xff1602b52bdc8a86, 0xffda9ccd49cad7e7,
    0xfffadd91a9b98cec, 0xffff61fa9ca0e619, 0xffffef2105923822,
    0xfffffe61be083611, 0xffffffdbf6da03df, 0xfffffffd2270a8ad,
    0xffffffffca5278cd, 0xfffffffffc5d60d0, 0xffffffffffc5618b,
    0xfffffffffffc8d07, 0xffffffffffffcf48, 0xfffffffffffffd78,
    0xffffffffffffffe0, 0xfffffffffffffffe, 0xffffffffffffffff,
];

/// Winners (positions 0..255) of block `block`, sub-stream `sub`, at `step`: 256 independent Bernoulli(2^-8) draws
/// realised with one call - the number of winners by inverse CDF, then distinct bytes of the block's byte stream
/// (words 2, 3 of the first call, then all four words of the calls with counter word 3 = 2, 3, ...).
pub fn lottery256(st: &Stream, block: u32, sub: u32, step: u32) -> Vec<u8> {
    let first = st.block(block, sub, step, 0);
    let u = ((first[0] as u64) << 32) | first[1] as u64;
    let k = LOTTERY_CDF_K8.iter().take_while(|&&t| u >= t).count();
    let (mut winners, mut seen, mut cur, mut word, mut byte, mut extra) = (Vec::new(), [false; 256], first, 2usize, 0u32, 0u32);
    while winners.len() < k {
        if word == 4 {
            cur = st.block(block, sub, step, 2 + extra);
            extra += 1;
            word = 0;
        }
        let b = (cur[word] >> (8 * byte)) as u8;
        byte += 1;
        if byte == 4 {
            byte = 0;
            word += 1;
        }
        if !seen[b as usize] {
            seen[b as usize] = true;
            winners.push(b);
        }
    }
    winners
}

/// examples/forest_fire.rs:360  `rng.random_bool(REGROWTH_PROBABILITY)` for an Empty entity
pub fn forest_regrow(seed: u64, replica: u32, uid: u32, step: u32, p: f64) -> bool {
    let st = Stream::new(seed, replica, b"FREG");
    if p * 256.0 <= 1.0 {
        // level 1: uid wins the lottery of its block of 256 uids (a real run would cache the block's winners)
        if !lottery256(&st, uid >> 8, 0, step).contains(&((uid & 255) as u8)) { return false; }
        bern(st.block(uid >> 2, 0, step, 1)[(uid & 3) as usize], threshold(p * 256.0))
    } else {
        bern(st.block(uid >> 2, 0, step, 1)[(uid & 3) as usize], threshold(p))
    }
}

/// examples/pandemic.rs:121-148  the three p = 1/2 draws: (stay, x_axis, negative)
pub fn pandemic_move(seed: u64, replica: u32, uid: u32, step: u32) -> (bool, bool, bool) {
    // one nibble per person: nibble uid & 7 of word (uid >> 3) & 3 of the call uid >> 5; draw k = bit 3-k clear
    let w = Stream::new(seed, replica, b"PMOV").block(uid >> 5, 0, step, 0)[((uid >> 3) & 3) as usize];
    let nib = (w >> (4 * (uid & 7))) & 0xf;
    (nib & 8 == 0, nib & 4 == 0, nib & 2 == 0)
}

/// examples/pandemic.rs:184  infect draw for the pair (healthy h, infected i)
pub fn pandemic_infect(seed: u64, replica: u32, h: u32, i: u32, step: u32, p: f64) -> bool {
    bern(Stream::new(seed, replica, b"PINF").block(h, i, step, 0)[0], threshold(p))
}
