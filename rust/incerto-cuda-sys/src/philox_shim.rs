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

/// examples/forest_fire.rs:360  `rng.random_bool(REGROWTH_PROBABILITY)` for an Empty entity
pub fn forest_regrow(seed: u64, replica: u32, uid: u32, step: u32, p: f64) -> bool {
    let st = Stream::new(seed, replica, b"FREG");
    if p * 256.0 <= 1.0 {
        let l1 = st.block(uid >> 4, 0, step, 0);
        if (l1[((uid >> 2) & 3) as usize] >> (8 * (uid & 3))) & 0xff != 0 { return false; }
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
