//! src/index/cuda.rs for the reference tree (v0.1.5): VectorIndex over the CUDA handles.  See INTEGRATION.md §2-§3c.
//! NOT compiled in the repo's build image (no cargo / rustc there).
use crate::{SearchResult, SimilarityMetric, Vector, VectorIndex};
use crate::errors::{VectorLiteError, VectorLiteResult};
use std::collections::HashMap;
use vectorlite_cuda_sys as sys;

fn metric_code(m: SimilarityMetric) -> i32 {      // declaration order of lib.rs:364-378
    match m { SimilarityMetric::Cosine => 0, SimilarityMetric::Euclidean => 1,
              SimilarityMetric::Manhattan => 2, SimilarityMetric::DotProduct => 3 }
}
fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::vl_last_error()).to_string_lossy().into_owned() }
}

/// Text + metadata stay on the host, exactly like `HNSWIndex::metadata` (hnsw.rs:78-82,210);
/// they are attached to the <= k hits only (the reference clones them for all n rows, flat.rs:111-112).
struct Meta { text: String, metadata: Option<serde_json::Value> }

pub struct CudaIndex { h: *mut sys::vl_index, dim: usize, meta: HashMap<u64, Meta>,
                       metric: Option<SimilarityMetric> }
// vl_index_search / get_vector / len are re-entrant on one handle; add / delete need exclusion —
// which is exactly what Arc<RwLock<VectorIndexWrapper>> (client.rs:245) provides.
unsafe impl Send for CudaIndex {}
unsafe impl Sync for CudaIndex {}
impl Drop for CudaIndex { fn drop(&mut self) { unsafe { sys::vl_index_destroy(self.h) } } }

impl CudaIndex {
    pub fn new_flat(dim: usize, device: i32) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        if unsafe { sys::vl_flat_create(dim as u32, device, &mut h) } != sys::VL_OK { return Err(last_error()); }
        Ok(Self { h, dim, meta: HashMap::new(), metric: None })
    }
    /// M/M0: 16/32 default, 8/16 memory-optimized, 32/64 high-accuracy (hnsw.rs:95-109) — runtime here.
    pub fn new_hnsw(dim: usize, metric: SimilarityMetric, m: u32, m0: u32, device: i32) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        if unsafe { sys::vl_hnsw_create(dim as u32, metric_code(metric), m, m0, 0, device, &mut h) } != sys::VL_OK {
            return Err(last_error());
        }
        Ok(Self { h, dim, meta: HashMap::new(), metric: Some(metric) })
    }
    pub fn metric(&self) -> Option<SimilarityMetric> { self.metric }
    /// Device beam = factor x ef, ef = min(k, len) as at hnsw.rs:437; 1 = equal ef, default 8.
    pub fn set_beam_factor(&mut self, factor: u32) -> Result<(), String> {
        if unsafe { sys::vl_hnsw_set_beam_factor(self.h, factor) } != sys::VL_OK { return Err(last_error()); }
        Ok(())
    }
    /// What `#[serde(skip)] index_internal` (hnsw.rs:199-200) leaves out of the .vlc: levels + adjacency of the
    /// graph.  `Collection::save_to_file` writes it to `<file>.graph` next to the JSON; the custom `Deserialize`
    /// (hnsw.rs:272-360) calls `restore_graph` instead of re-inserting every vector when that file is present and
    /// valid.  Err when the graph holds soft-deleted nodes (rebuild on load then, as the reference does).
    pub fn graph_blob(&self) -> Result<Vec<u8>, String> {
        let mut n = 0u64;
        if unsafe { sys::vl_hnsw_graph_bytes(self.h, &mut n) } != sys::VL_OK { return Err(last_error()); }
        let (mut buf, mut written) = (vec![0u8; n as usize], 0u64);
        if unsafe { sys::vl_hnsw_export_graph(self.h, buf.as_mut_ptr().cast(), n, &mut written) } != sys::VL_OK {
            return Err(last_error());
        }
        buf.truncate(written as usize);
        Ok(buf)
    }
    /// `ids` / `rows` in the insertion order the blob was exported with (vl_index_export); the index must be empty.
    pub fn restore_graph(&mut self, ids: &[u64], rows: &[f32], blob: &[u8]) -> Result<(), String> {
        match unsafe { sys::vl_hnsw_import_graph(self.h, ids.as_ptr(), rows.as_ptr(), ids.len() as u64,
                                                 blob.as_ptr().cast(), blob.len() as u64) } {
            sys::VL_OK => Ok(()),
            _ => Err(last_error()),
        }
    }
    pub fn max_id(&self) -> Option<u64> {                       // flat.rs:76-78, hnsw.rs:267-269
        let mut id = 0u64;
        (unsafe { sys::vl_index_max_id(self.h, &mut id) } == sys::VL_OK).then_some(id)
    }
}

impl VectorIndex for CudaIndex {
    fn add(&mut self, v: Vector) -> Result<(), String> {
        // the status codes map back onto the strings client.rs:334-345,366-377 substring-match
        match unsafe { sys::vl_index_add_f64(self.h, v.id, v.values.as_ptr(), v.values.len() as u32) } {
            sys::VL_OK => { self.meta.insert(v.id, Meta { text: v.text, metadata: v.metadata }); Ok(()) }
            _ => Err(last_error()),   // "Vector dimension mismatch…" / "Vector ID {} already exists"
        }
    }
    fn delete(&mut self, id: u64) -> Result<(), String> {
        match unsafe { sys::vl_index_delete(self.h, id) } {
            sys::VL_OK => { self.meta.remove(&id); Ok(()) }
            _ => Err(last_error()),   // HNSW only: "Vector ID {} does not exist" (flat: missing id is Ok)
        }
    }
    fn search(&self, q: &[f64], k: usize, m: SimilarityMetric) -> VectorLiteResult<Vec<SearchResult>> {
        let (mut ids, mut scores, mut count) = (vec![0u64; k], vec![0f64; k], 0u32);
        let st = unsafe { sys::vl_index_search_f64(self.h, q.as_ptr(), 1, q.len() as u32, k as u32,
                     metric_code(m), 0 /* ef = min(k,len), hnsw.rs:437 */,
                     ids.as_mut_ptr(), scores.as_mut_ptr(), &mut count) };
        match st {
            sys::VL_OK => Ok((0..count as usize).map(|i| {
                let meta = self.meta.get(&ids[i]);
                SearchResult { id: ids[i], score: scores[i],
                               text: meta.map(|m| m.text.clone()).unwrap_or_default(),
                               metadata: meta.and_then(|m| m.metadata.clone()) }
            }).collect()),
            sys::VL_ERR_DIM => Err(VectorLiteError::DimensionMismatch { expected: self.dim, actual: q.len() }),
            sys::VL_ERR_METRIC_MISMATCH => Err(VectorLiteError::MetricMismatch { requested: m, index: self.metric.unwrap() }),
            sys::VL_ERR_NAN => panic!("NaN similarity"),   // what flat.rs:116 does
            _ => Err(VectorLiteError::InternalError(last_error())),
        }
    }
    fn len(&self) -> usize { unsafe { sys::vl_index_len(self.h) as usize } }
    fn is_empty(&self) -> bool { self.len() == 0 }
    fn get_vector(&self, id: u64) -> Option<Vector> {
        let mut v = vec![0f32; self.dim];
        (unsafe { sys::vl_index_get_vector(self.h, id, v.as_mut_ptr()) } == sys::VL_OK).then(|| {
            let meta = self.meta.get(&id);
            Vector { id, values: v.into_iter().map(f64::from).collect(),
                     text: meta.map(|m| m.text.clone()).unwrap_or_default(),
                     metadata: meta.and_then(|m| m.metadata.clone()) }
        })
    }
    fn dimension(&self) -> usize { self.dim }
}

// ---- several GPUs in the one server process ----
/*
pub struct ShardedCudaFlat {
    shards: Vec<CudaIndex>,            // one per device; shard g holds storage positions [base_g, base_g + n_g)
    group: *mut vl_group,              // borrows the handles: dropped first
    where_: HashMap<u64, usize>,       // id -> shard (duplicate-id check over the WHOLE store, delete routing)
    tail: usize, shard_rows: usize,    // appends go to the shard that owns the tail of the storage order
}
// add:    dim / duplicate checks, then shards[tail].add(v); tail moves on when the shard holds shard_rows rows
// delete: shards[where_[id]].delete(id)            (order-preserving inside the shard; missing id = Ok)
// search: vl_group_search                          (all shards scan at once; stable merge in shard order ==
//                                                   flat.rs:116 over the whole store)
// load:   rows split evenly, one vl_index_add_batch per shard
*/
