//! Raw bindings to libvectorlite_cuda.so (include/vectorlite_cuda.h in the vectorlite-b200 repo).
//! NOT compiled in the repo's build image (no cargo / rustc there): the same calls, in the same order, are made by
//! vectorlite_b200/__init__.py (ctypes) and include/vectorlite.hpp (C++), which the GPU tests exercise.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct vl_index { _private: [u8; 0] }

pub const VL_OK: c_int = 0;
pub const VL_ERR_DIM: c_int = 1;             // flat.rs:84 / hnsw.rs:365 / DimensionMismatch
pub const VL_ERR_DUP_ID: c_int = 2;          // flat.rs:87 / hnsw.rs:369
pub const VL_ERR_NOT_FOUND: c_int = 3;       // hnsw.rs:402
pub const VL_ERR_METRIC_MISMATCH: c_int = 4; // hnsw.rs:426-429
pub const VL_ERR_NAN: c_int = 8;             // the reference panics at flat.rs:116

#[link(name = "vectorlite_cuda")]
extern "C" {
    pub fn vl_flat_create(dim: u32, device: c_int, out: *mut *mut vl_index) -> c_int;
    pub fn vl_hnsw_create(dim: u32, metric: c_int, m: u32, m0: u32, ef_construction: u32,
                          device: c_int, out: *mut *mut vl_index) -> c_int;
    pub fn vl_index_destroy(h: *mut vl_index);
    pub fn vl_index_add_f64(h: *mut vl_index, id: u64, values: *const f64, len: u32) -> c_int;
    pub fn vl_index_add_batch(h: *mut vl_index, ids: *const u64, rows: *const f32, n: u64) -> c_int;
    pub fn vl_index_delete(h: *mut vl_index, id: u64) -> c_int;
    pub fn vl_index_search_f64(h: *mut vl_index, queries: *const f64, nq: u32, qdim: u32, k: u32,
                               metric: c_int, ef: u32, out_ids: *mut u64, out_scores: *mut f64,
                               out_counts: *mut u32) -> c_int;
    pub fn vl_index_len(h: *const vl_index) -> u64;
    pub fn vl_index_dim(h: *const vl_index) -> u32;
    pub fn vl_index_metric(h: *const vl_index) -> c_int;
    pub fn vl_index_max_id(h: *const vl_index, out_id: *mut u64) -> c_int;
    pub fn vl_index_get_vector(h: *const vl_index, id: u64, out_values: *mut f32) -> c_int;
    pub fn vl_index_export(h: *const vl_index, first: u64, cap: u64, out_ids: *mut u64,
                           out_rows: *mut f32, out_n: *mut u64) -> c_int;
    // HNSW only: where bulk loads build the graph (0 auto, 1 host, 2 device) and which score is returned
    // (0 = exact Flat similarity, 1 = the reference's quantised score, hnsw.rs:478 + 51-75, bit for bit)
    pub fn vl_hnsw_set_builder(h: *mut vl_index, builder: c_int) -> c_int;
    pub fn vl_hnsw_set_score_mode(h: *mut vl_index, mode: c_int) -> c_int;
    pub fn vl_hnsw_set_beam_factor(h: *mut vl_index, factor: u32) -> c_int;
    pub fn vl_hnsw_graph_bytes(h: *const vl_index, out_bytes: *mut u64) -> c_int;
    pub fn vl_hnsw_export_graph(h: *const vl_index, buf: *mut c_void, cap: u64, out_written: *mut u64) -> c_int;
    pub fn vl_hnsw_import_graph(h: *mut vl_index, ids: *const u64, rows: *const f32, n: u64, blob: *const c_void, bytes: u64) -> c_int;
    pub fn vl_last_error() -> *const c_char;
}

// ---- shard group: one flat handle per GPU of this process, searched as one index (csrc/group.cpp) ----
#[repr(C)] pub struct vl_group { _private: [u8; 0] }

#[link(name = "vectorlite_cuda")]
extern "C" {
    pub fn vl_group_create(shards: *const *mut vl_index, n: u32, out: *mut *mut vl_group) -> c_int;
    pub fn vl_group_destroy(g: *mut vl_group);
    pub fn vl_group_search(g: *mut vl_group, queries: *const f32, nq: u32, qdim: u32, k: u32, metric: c_int,
                           out_ids: *mut u64, out_scores: *mut f64, out_counts: *mut u32) -> c_int;
}
