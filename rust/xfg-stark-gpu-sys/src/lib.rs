//! Raw bindings, one item per declaration of `include/xfg_stark.h` (kept in the same order).
//! The product crate `xfg-stark` has `#![deny(unsafe_code)]` (src/lib.rs:17), hence this separate `-sys` crate.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const XFG_NUM_PUB_INPUTS: usize = 12;
pub const XFG_NUM_STAGES: usize = 9;
pub const XFG_EXT_NONE: u32 = 1;
pub const XFG_EXT_QUADRATIC: u32 = 2;

pub const XFG_OK: c_int = 0;
pub const XFG_ERR_BAD_ARGS: c_int = 1;
pub const XFG_ERR_BAD_OPTIONS: c_int = 2;
pub const XFG_ERR_UNSUPPORTED_OPTIONS: c_int = 3;
pub const XFG_ERR_UNSUPPORTED_EXTENSION: c_int = 4;
pub const XFG_ERR_UNSATISFIED_CONSTRAINT: c_int = 5;
pub const XFG_ERR_BUFFER_TOO_SMALL: c_int = 6;
pub const XFG_ERR_CUDA: c_int = 7;
pub const XFG_ERR_INVALID_INPUT: c_int = 8;
pub const XFG_ERR_TOO_LARGE: c_int = 9;

#[repr(C)]
pub struct xfg_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct xfg_wide { _private: [u8; 0] }

// ---- generic AIR front-end (xfg_air_desc) ----
pub const XFG_OP_ADD: u32 = 0;
pub const XFG_OP_SUB: u32 = 1;
pub const XFG_OP_MUL: u32 = 2;
pub const XFG_AIR_MAX_WIDTH: u32 = 128;
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct xfg_air_instr { pub op: u32, pub a: u32, pub b: u32 }
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct xfg_assertion { pub column: u32, pub step: u32, pub value: u64 }
#[repr(C)]
pub struct xfg_air_desc {
    pub width: u32,
    pub num_pub_inputs: u32,
    pub num_constants: u32,
    pub num_instr: u32,
    pub num_constraints: u32,
    pub num_assertions: u32,
    pub pub_inputs: *const u64,
    pub constants: *const u64,
    pub code: *const xfg_air_instr,
    pub constraint_values: *const u32,
    pub assertions: *const xfg_assertion,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct xfg_options {
    pub num_queries: u32,
    pub blowup_factor: u32,
    pub grinding_factor: u32,
    pub field_extension: u32,
    pub fri_folding_factor: u32,
    pub fri_remainder_max_degree: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct xfg_air_consts {
    pub pub_inputs: [u64; XFG_NUM_PUB_INPUTS],
    pub txn_hash: u64,
    pub recipient_hash: u64,
    pub nullifier: u64,
    pub commitment: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct xfg_stage_times {
    pub stage_ms: [f32; XFG_NUM_STAGES],
    pub h2d_ms: f32,
    pub device_ms: f32,
    pub total_ms: f32,
    pub kernel_launches: u32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct xfg_verify_times {
    pub host_parse_ms: f32,
    pub h2d_ms: f32,
    pub kernel_ms: f32,
    pub total_ms: f32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}
pub const XFG_VERIFY_OK: i32 = 0;
pub const XFG_FORM_CANONICAL: u32 = 0;
pub const XFG_FORM_MONTGOMERY: u32 = 1;

extern "C" {
    pub fn xfg_create(device: c_int, max_n_log2: u32, num_slots: u32, out: *mut *mut xfg_ctx) -> c_int;
    pub fn xfg_destroy(ctx: *mut xfg_ctx);
    pub fn xfg_strerror(code: c_int) -> *const c_char;
    pub fn xfg_last_error(ctx: *const xfg_ctx) -> *const c_char;
    pub fn xfg_set_profiling(ctx: *mut xfg_ctx, on: c_int) -> c_int;
    pub fn xfg_get_profile(ctx: *mut xfg_ctx, cap: u32, count: *mut u32, names: *mut *const c_char, ms: *mut f32, launches: *mut u32) -> c_int;
    pub fn xfg_prove_burn_mint(ctx: *mut xfg_ctx, trace_colmajor: *const u64, n_log2: u32, air: *const xfg_air_consts, options: *const xfg_options,
                               out: *mut u8, out_cap: usize, out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    pub fn xfg_prove_burn_mint_device(ctx: *mut xfg_ctx, d_trace_colmajor: *const u64, n_log2: u32, air: *const xfg_air_consts,
                                      options: *const xfg_options, out: *mut u8, out_cap: usize, out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    /// cols: seven column pointers (`TraceTable::get_column(i).as_ptr()`); form: XFG_FORM_CANONICAL / XFG_FORM_MONTGOMERY
    pub fn xfg_prove_burn_mint_cols(ctx: *mut xfg_ctx, cols: *const *const u64, form: u32, n_log2: u32, air: *const xfg_air_consts,
                                    options: *const xfg_options, out: *mut u8, out_cap: usize, out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    /// debug: workspace guard zones (memory-safety check where compute-sanitizer is unavailable)
    pub fn xfg_debug_guard_fill(ctx: *mut xfg_ctx) -> c_int;
    pub fn xfg_debug_guard_check(ctx: *mut xfg_ctx, n_log2: u32, field_extension: u32, width: u32, fri_remainder_max_degree: u32, violations: *mut u64, first_region: *mut i32) -> c_int;
    pub fn xfg_debug_poke_guard(ctx: *mut xfg_ctx, n_log2: u32) -> c_int;
    pub fn xfg_host_register(ctx: *mut xfg_ctx, ptr: *const c_void, bytes: usize) -> c_int;
    pub fn xfg_host_unregister(ctx: *mut xfg_ctx, ptr: *const c_void) -> c_int;
    pub fn xfg_prove_burn_mint_batch(ctx: *mut xfg_ctx, count: u32, traces: *const *const u64, n_log2: u32, airs: *const xfg_air_consts,
                                     options: *const xfg_options, out: *mut u8, out_stride: usize, out_lens: *mut usize, total_ms: *mut f32) -> c_int;
    pub fn xfg_burn_mint_pack_inputs(ctx: *mut xfg_ctx, burn_amount: u64, mint_amount: u64, tx_prefix_hash: *const u8, recipient_address: *const u8,
                                     recipient_len: usize, secret: *const u8, secret_len: usize, network_id: u32, target_chain_id: u32,
                                     commitment_version: u32, out: *mut xfg_air_consts) -> c_int;
    pub fn xfg_burn_mint_build_trace(air: *const xfg_air_consts, n_log2: u32, trace_colmajor_out: *mut u64) -> c_int;
    pub fn xfg_prove_burn_mint_from_inputs(ctx: *mut xfg_ctx, burn_amount: u64, mint_amount: u64, tx_prefix_hash: *const u8, recipient_address: *const u8,
                                           recipient_len: usize, secret: *const u8, secret_len: usize, network_id: u32, target_chain_id: u32,
                                           commitment_version: u32, n_log2: u32, options: *const xfg_options, out: *mut u8, out_cap: usize,
                                           out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    pub fn xfg_create_ex(device: c_int, max_n_log2: u32, num_slots: u32, max_width: u32, out: *mut *mut xfg_ctx) -> c_int;
    pub fn xfg_prove_air(ctx: *mut xfg_ctx, air: *const xfg_air_desc, trace_colmajor: *const u64, n_log2: u32, options: *const xfg_options, out: *mut u8,
                         out_cap: usize, out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    pub fn xfg_prove_air_device(ctx: *mut xfg_ctx, air: *const xfg_air_desc, d_trace_colmajor: *const u64, n_log2: u32, options: *const xfg_options,
                                out: *mut u8, out_cap: usize, out_len: *mut usize, times: *mut xfg_stage_times) -> c_int;
    pub fn xfg_prove_air_batch(ctx: *mut xfg_ctx, count: u32, airs: *const xfg_air_desc, traces: *const *const u64, n_log2: u32, options: *const xfg_options,
                               out: *mut u8, out_stride: usize, out_lens: *mut usize, total_ms: *mut f32) -> c_int;
    pub fn xfg_air_compile_check(air: *const xfg_air_desc, n_log2: u32, num_instr: *mut u32, num_slots: *mut u32, num_groups: *mut u32,
                                 cur: *const u64, next: *const u64, out: *mut u64) -> c_int;
    pub fn xfg_verify_burn_mint_batch(ctx: *mut xfg_ctx, count: u32, proofs: *const *const u8, proof_lens: *const usize, air: *const xfg_air_consts,
                                      acceptable: *const xfg_options, results: *mut i32, times: *mut xfg_verify_times) -> c_int;
    pub fn xfg_verify_air_batch(ctx: *mut xfg_ctx, count: u32, proofs: *const *const u8, proof_lens: *const usize, airs: *const xfg_air_desc,
                                acceptable: *const xfg_options, results: *mut i32, times: *mut xfg_verify_times) -> c_int;
    pub fn xfg_verify_strerror(code: c_int) -> *const c_char;
    pub fn xfg_ntt(ctx: *mut xfg_ctx, data: *mut u64, n_log2: u32, batch: u32, inverse: c_int) -> c_int;
    pub fn xfg_lde_commit(ctx: *mut xfg_ctx, cols_colmajor: *const u64, n_log2: u32, cols: u32, lde_out: *mut u64, root_out: *mut u8) -> c_int;
    pub fn xfg_merkle_root(ctx: *mut xfg_ctx, leaves: *const u8, count: usize, root_out: *mut u8, nodes_out: *mut u8) -> c_int;
    pub fn xfg_eval_constraints(ctx: *mut xfg_ctx, lde: *const u64, n_log2: u32, air: *const xfg_air_consts, ext: u32, coeffs: *const u64, out: *mut u64) -> c_int;
    pub fn xfg_fri_fold_layer(ctx: *mut xfg_ctx, evals: *const u64, nl_log2: u32, ext: u32, alpha: *const u64, out: *mut u64) -> c_int;
    pub fn xfg_field_selftest(ctx: *mut xfg_ctx, op: u32, a: *const u64, b: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn xfg_int_pipe_peak(ctx: *mut xfg_ctx, gops: *mut f64) -> c_int;
    pub fn xfg_set_graphs(ctx: *mut xfg_ctx, on: c_int) -> c_int;
    // one wide trace sharded over the GPUs of a box (BASELINE config 5)
    pub fn xfg_wide_create(ctx: *mut xfg_ctx, n_log2: u32, total_cols: u32, num_ranks: u32, rank: u32, out: *mut *mut xfg_wide) -> c_int;
    pub fn xfg_wide_destroy(w: *mut xfg_wide);
    pub fn xfg_wide_recv_ptr(w: *mut xfg_wide) -> *mut std::os::raw::c_void;
    pub fn xfg_wide_ipc_handle(w: *mut xfg_wide, out: *mut u8) -> c_int;
    pub fn xfg_wide_open_peers(w: *mut xfg_wide, handles: *const u8) -> c_int;
    pub fn xfg_wide_set_peer_ptrs(w: *mut xfg_wide, ptrs: *const *mut std::os::raw::c_void) -> c_int;
    pub fn xfg_wide_extend(w: *mut xfg_wide, d_cols_local: *const u64, device_ms: *mut f32) -> c_int;
    pub fn xfg_wide_commit(w: *mut xfg_wide, subtree_root: *mut u8, device_ms: *mut f32) -> c_int;
    pub fn xfg_wide_read_recv(w: *mut xfg_wide, out: *mut u64) -> c_int;
    pub fn xfg_pipe_probe(ctx: *mut xfg_ctx, mode: c_int, gops: *mut f64) -> c_int;
    pub fn xfg_hash_rows(ctx: *mut xfg_ctx, rows_rowmajor: *const u64, rows: usize, limbs: u32, out: *mut u8) -> c_int;
}
