// bindings/ffi.rs — the Rust side of the drop-in boundary: copy to richox/orz `src/ffi.rs`, uncomment `// pub mod ffi;` at
// src/lib.rs:10 and import `ffi::{LZDecoder, LZEncoder}` instead of `lz::{LZDecoder, LZEncoder}` at src/lib.rs:26-29.
// `orz::encode` / `orz::decode` (lib.rs:58-129) and the CLI (main.rs) then run unchanged on the B200: window buffers, chunk loop,
// framing and slide stay in Rust; every LZEncoder::encode / LZDecoder::decode call (lib.rs:78,116) crosses this boundary.
//
// Declarations mirror include/orz_b200.h one to one.  Cannot be compiled in this repository's containers (no Rust toolchain);
// tools/make_ref_golden.sh --with-ffi builds it into the reference crate wherever nightly Rust exists:
//     RUSTFLAGS="-L /path/to/orz_b200 -l orz_b200" cargo +nightly build --release
use std::io::{Error, ErrorKind, Result};
use crate::lz::LZCfg;

#[repr(C)] pub struct OrzLzEnc { _p: [u8; 0] }
#[repr(C)] pub struct OrzLzDec { _p: [u8; 0] }

extern "C" {
    fn orz_lzenc_new(device: i32) -> *mut OrzLzEnc;
    fn orz_lzenc_free(e: *mut OrzLzEnc);
    fn orz_lzenc_forward(e: *mut OrzLzEnc, forward_len: usize) -> i32;
    fn orz_lzenc_encode(e: *mut OrzLzEnc, cfg: *const LZCfg, sbuf: *const u8, sbuf_len: usize,
                        tbuf: *mut u8, tbuf_cap: usize, spos: usize,
                        spos_out: *mut usize, tlen_out: *mut usize) -> i32;
    fn orz_lzdec_new(device: i32) -> *mut OrzLzDec;
    fn orz_lzdec_free(d: *mut OrzLzDec);
    fn orz_lzdec_forward(d: *mut OrzLzDec, forward_len: usize) -> i32;
    fn orz_lzdec_decode(d: *mut OrzLzDec, tbuf: *const u8, tlen: usize, sbuf: *mut u8, sbuf_cap: usize,
                        spos: usize, spos_end: *mut usize) -> i32;
}

/// Same surface as lz::LZEncoder (lz.rs:69-95), so lib.rs:64,78,84 compile unchanged.
pub struct LZEncoder(*mut OrzLzEnc);
impl LZEncoder {
    pub fn new() -> Self {
        let p = unsafe { orz_lzenc_new(0) };
        assert!(!p.is_null(), "orz_b200: no sm_100 device (there is no CPU fallback)");
        LZEncoder(p)
    }
    pub fn forward(&mut self, forward_len: usize) { unsafe { orz_lzenc_forward(self.0, forward_len); } }
    pub fn encode(&mut self, cfg: &LZCfg, sbuf: &[u8], tbuf: &mut [u8], spos: usize) -> (usize, usize) {
        let (mut s, mut t) = (0usize, 0usize);
        // sbuf is a sub-slice of sbvec_buf with SBVEC_SENTINEL_LEN bytes either side (lib.rs:67-69,77)
        let rc = unsafe { orz_lzenc_encode(self.0, cfg, sbuf.as_ptr(), sbuf.len(), tbuf.as_mut_ptr(), tbuf.len(), spos, &mut s, &mut t) };
        assert_eq!(rc, 0, "orz_b200 encode failed: {rc}");
        (s, t)
    }
}
impl Drop for LZEncoder { fn drop(&mut self) { unsafe { orz_lzenc_free(self.0) } } }

/// Same surface as lz::LZDecoder (lz.rs:348-366), so lib.rs:99,116,122 compile unchanged.
pub struct LZDecoder(*mut OrzLzDec);
impl LZDecoder {
    pub fn new() -> Self { let p = unsafe { orz_lzdec_new(0) }; assert!(!p.is_null()); LZDecoder(p) }
    pub fn forward(&mut self, forward_len: usize) { unsafe { orz_lzdec_forward(self.0, forward_len); } }
    pub fn decode(&mut self, tbuf: &[u8], sbuf: &mut [u8], spos: usize) -> Result<usize> {
        let mut end = 0usize;
        match unsafe { orz_lzdec_decode(self.0, tbuf.as_ptr(), tbuf.len(), sbuf.as_mut_ptr(), sbuf.len(), spos, &mut end) } {
            0 => Ok(end),
            -1 => Err(Error::from(ErrorKind::InvalidData)),
            -3 => Err(Error::from(ErrorKind::UnexpectedEof)),
            rc => Err(Error::new(ErrorKind::Other, format!("orz_b200: {rc}"))),
        }
    }
}
impl Drop for LZDecoder { fn drop(&mut self) { unsafe { orz_lzdec_free(self.0) } } }
