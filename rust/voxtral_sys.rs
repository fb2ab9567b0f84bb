// rust/voxtral_sys.rs -- raw `extern "C"` bindings for include/voxtral.h (libvoxtral_b200.so).
//
// SOURCE ONLY: there is no Rust toolchain in the build environment, so this file has never been compiled;
// it is the `src/b200/sys.rs` a maintainer of TrevorS/voxtral-mini-realtime-rs would add behind a `b200` cargo feature
// (Cargo.toml: `b200 = []`; build.rs: `println!("cargo:rustc-link-lib=dylib=voxtral_b200")`).  INTEGRATION.md shows the
// safe wrapper (`mod.rs`) and the call-site changes in src/bin/transcribe.rs.  tests/test_host_abi.py checks that every
// function declared here is exported by the library with the same name.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct vox_gguf { _p: [u8; 0] }
#[repr(C)] pub struct vox_model { _p: [u8; 0] }
#[repr(C)] pub struct vox_session { _p: [u8; 0] }
#[repr(C)] pub struct vox_q4 { _p: [u8; 0] }
#[repr(C)] pub struct vox_mel { _p: [u8; 0] }
#[repr(C)] pub struct vox_tokenizer { _p: [u8; 0] }
#[repr(C)] pub struct vox_stream_pool { _p: [u8; 0] }

#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct vox_timings {
    pub preprocess_ms: f32, pub encode_ms: f32, pub decode_ms: f32, pub total_ms: f32,
    pub prefill_ms: f32, pub decode_tokens: i32, pub seq_len: i32,
}
#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct vox_stream_stats {
    pub gpu_ms: f32, pub live_sessions: i32, pub mel_frames: i32, pub encoder_rows: i32,
    pub prefills: i32, pub decode_steps: i32, pub decode_rows: i32,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct vox_pad_config { pub sample_rate: u32, pub n_left_pad_tokens: u32, pub frame_rate: f32,
                            pub extra_right_pad_tokens: u32 }

extern "C" {
    pub fn vox_last_error() -> *const c_char;
    // src/gguf/reader.rs + loader.rs
    pub fn vox_gguf_open(path: *const c_char, out: *mut *mut vox_gguf) -> i32;
    pub fn vox_gguf_open_shards(bufs: *const *const c_void, lens: *const usize, n: usize,
                                out: *mut *mut vox_gguf) -> i32;
    pub fn vox_gguf_close(g: *mut vox_gguf);
    pub fn vox_model_load_gguf(path: *const c_char, device: i32, out: *mut *mut vox_model) -> i32;
    pub fn vox_model_load_gguf_handle(g: *mut vox_gguf, device: i32, out: *mut *mut vox_model) -> i32;
    pub fn vox_model_free(m: *mut vox_model);
    // src/gguf/model.rs
    pub fn vox_session_create(m: *mut vox_model, max_batch: i32, max_mel_frames: i32,
                              out: *mut *mut vox_session) -> i32;
    pub fn vox_session_set_delay(s: *mut vox_session, delay_tokens: f32) -> i32;
    pub fn vox_encode_audio(s: *mut vox_session, mel: *const f32, b: i32, t: i32,
                            audio_embeds: *mut f32, cap: usize, seq_len: *mut i32) -> i32;
    pub fn vox_transcribe_streaming(s: *mut vox_session, mel: *const f32, b: i32, t: i32,
                                    out_ids: *mut i32, cap: usize, n_out: *mut i32,
                                    tm: *mut vox_timings) -> i32;
    pub fn vox_transcribe_pcm(s: *mut vox_session, samples: *const f32, b: i32, n: usize,
                              peak_normalize: i32, out_ids: *mut i32, cap: usize, n_out: *mut i32,
                              tm: *mut vox_timings) -> i32;
    pub fn vox_generate_step_with_cache(s: *mut vox_session, ids: *const i32, b: i32, m: i32,
                                        logits: *mut f32, cap: usize) -> i32;
    // forward_streaming (model.rs:801-814) and the device-side incremental decode (model.rs:857-867 without the
    // logits round trip)
    pub fn vox_forward_streaming(s: *mut vox_session, mel: *const f32, b: i32, t: i32, ids: *const i32, n_ids: i32,
                                 logits: *mut f32, cap: usize) -> i32;
    pub fn vox_prefill(s: *mut vox_session, ids: *const i32, b: i32, m: i32, add_audio: i32, next_tok: *mut i32) -> i32;
    pub fn vox_decode_step(s: *mut vox_session, tok: *const i32, b: i32, add_audio: i32, next_tok: *mut i32) -> i32;
    pub fn vox_session_reset(s: *mut vox_session) -> i32;
    pub fn vox_session_free(s: *mut vox_session);
    // src/gguf/{tensor,linear,op}.rs
    pub fn vox_q4_tensor_create(bytes: *const u8, nbytes: usize, n: i64, k: i64, device: i32,
                                out: *mut *mut vox_q4) -> i32;
    pub fn vox_q4_matmul(w: *const vox_q4, x_dev: *const f32, y_dev: *mut f32, b: i32, m: i32,
                         bias_dev: *const f32, stream: *mut c_void) -> i32;
    pub fn vox_q4_matmul_host(w: *const vox_q4, x: *const f32, y: *mut f32, b: i32, m: i32,
                              bias: *const f32) -> i32;
    pub fn vox_q4_tensor_free(w: *mut vox_q4);
    // src/audio/{mel,pad,io}.rs, src/models/time_embedding.rs
    pub fn vox_mel_create(device: i32, out: *mut *mut vox_mel) -> i32;
    pub fn vox_mel_num_frames(n: usize) -> usize;
    pub fn vox_mel_compute_log(m: *mut vox_mel, samples: *const f32, n: usize, out: *mut f32, cap: usize) -> i32;
    pub fn vox_mel_free(m: *mut vox_mel);
    pub fn vox_peak_normalize(samples: *mut f32, n: usize, target: f32) -> i32;
    pub fn vox_pad_audio_len(n: usize, cfg: *const vox_pad_config) -> usize;
    pub fn vox_pad_audio(inp: *const f32, n: usize, cfg: *const vox_pad_config, out: *mut f32,
                         cap: usize, out_len: *mut usize) -> i32;
    pub fn vox_time_embedding(t: f32, dim: i32, out: *mut f32) -> i32;
    pub fn vox_stream_progress(n_samples: usize, ended: i32, reshape_factor: i32, prefix_len: i32, out: *mut i64) -> i32;
    // streaming sessions: Q4AudioEncoder::forward_with_cache (model.rs:437-452), encode_audio_with_cache (790-799)
    pub fn vox_stream_pool_create(m: *mut vox_model, max_sessions: i32, max_seconds: f32, out: *mut *mut vox_stream_pool) -> i32;
    pub fn vox_stream_open(p: *mut vox_stream_pool, session: *mut i32) -> i32;
    pub fn vox_stream_push_pcm(p: *mut vox_stream_pool, session: i32, samples: *const f32, n: usize) -> i32;
    pub fn vox_stream_finish(p: *mut vox_stream_pool, session: i32) -> i32;
    pub fn vox_stream_tick(p: *mut vox_stream_pool, stats: *mut vox_stream_stats) -> i32;
    pub fn vox_stream_poll_ids(p: *mut vox_stream_pool, session: i32, ids: *mut i32, cap: usize, n: *mut usize, done: *mut i32) -> i32;
    pub fn vox_stream_encode_chunk(p: *mut vox_stream_pool, session: i32, mel: *const f32, t_frames: i32, audio_embeds: *mut f32,
                                   cap: usize, n: *mut i32) -> i32;
    pub fn vox_stream_close(p: *mut vox_stream_pool, session: i32) -> i32;
    pub fn vox_stream_pool_free(p: *mut vox_stream_pool);
    // src/tokenizer/mod.rs
    pub fn vox_tokenizer_from_file(path: *const c_char, out: *mut *mut vox_tokenizer) -> i32;
    pub fn vox_tokenizer_decode(t: *const vox_tokenizer, ids: *const u32, n: usize, buf: *mut c_char,
                                cap: usize, written: *mut usize) -> i32;
    pub fn vox_tokenizer_free(t: *mut vox_tokenizer);
}
