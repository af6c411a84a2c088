//! `KGraph -> ANNKGCSR` hand-off (SOURCE ONLY; goes into annembed as src/fromhnsw/kgraph_csr.rs behind `--features cuda`).
//!
//! The reference's KGraph has no serialisation (only the Hnsw is dumpable, in hnsw_rs's private format).  This module
//! flattens `KGraph::get_neighbours()` (src/fromhnsw/kgraph.rs:157: `&Vec<Vec<OutEdge<F>>>`, rows sorted by increasing
//! distance, :508-509) and the idx -> DataId map (`get_data_id_from_idx`, :335-337) into the CSR arrays the C ABI
//! takes (include/annembed_cuda.h `annembed_cuda_set_graph_csr`), and writes / reads the interchange file shared with
//! the Python harness (annembed_b200/kgraph.py `write_csr` / `read_csr`):
//!
//!     magic  8 bytes  b"ANNKGCSR"
//!     u32 version (=1), u32 flags (=0)
//!     u64 n, u64 E, u64 max_nbng
//!     u64 row_ptr[n+1] ; u32 col[E] ; f32 dist[E] ; u64 data_id[n]          (little endian)
//!
//! "graph built once by hnsw_rs, serialised, fed identically to both implementations" (north star) = `write_csr` after
//! `kgraph_from_hnsw_all` (:440-579) in the reference process, `read_csr` in every consumer.
use std::fs::File;
use std::io::{BufReader, BufWriter, Read, Write};
use std::path::Path;

use num_traits::{Float, FromPrimitive};

use crate::fromhnsw::kgraph::KGraph;

pub const MAGIC: &[u8; 8] = b"ANNKGCSR";
pub const VERSION: u32 = 1;

/// Flat CSR view of a KGraph: what `annembed_cuda_set_graph_csr` consumes.
#[derive(Clone, Debug, Default, PartialEq)]
pub struct KGraphCsr {
    pub max_nbng: u64,
    pub row_ptr: Vec<u64>, // n + 1
    pub col: Vec<u32>,     // E, node indices (rank in the graph, not DataId)
    pub dist: Vec<f32>,    // E, ascending inside each row
    pub data_id: Vec<u64>, // n, DataId of node index i (kgraph.rs:335)
}

impl KGraphCsr {
    pub fn nb_nodes(&self) -> usize {
        self.row_ptr.len().saturating_sub(1)
    }
    pub fn nb_edges(&self) -> usize {
        self.col.len()
    }

    /// ≙ reading `kgraph.get_neighbours()` (kgraph.rs:157) row by row.  Fails like to_proba_edges does
    /// (src/tools/kdumap.rs:75-85 exits the process there) when a node has no neighbour, and when the graph does not
    /// fit the device limits (n, E < 2^32 - 1).
    pub fn from_kgraph<F>(kgraph: &KGraph<F>) -> Result<Self, String>
    where
        F: FromPrimitive + Float + std::fmt::UpperExp + Sync + Send + std::iter::Sum,
    {
        let neighbours = kgraph.get_neighbours();
        let n = kgraph.get_nb_nodes();
        if neighbours.len() != n {
            return Err(format!("KGraph: {} neighbour lists for {} nodes", neighbours.len(), n));
        }
        let nb_edges: usize = neighbours.iter().map(|v| v.len()).sum();
        if n as u64 >= u32::MAX as u64 || nb_edges as u64 >= u32::MAX as u64 {
            return Err("KGraph too large for the device path (n, E must be < 2^32 - 1)".to_string());
        }
        let mut csr = KGraphCsr {
            max_nbng: kgraph.get_max_nbng() as u64,
            row_ptr: Vec::with_capacity(n + 1),
            col: Vec::with_capacity(nb_edges),
            dist: Vec::with_capacity(nb_edges),
            data_id: Vec::with_capacity(n),
        };
        csr.row_ptr.push(0);
        for (i, edges) in neighbours.iter().enumerate() {
            if edges.is_empty() {
                return Err(format!("to_proba_edges: node rank {} has no neighbour", i)); // kdumap.rs:75-85
            }
            let mut prev = f32::NEG_INFINITY;
            for e in edges {
                let w = e.weight.to_f32().unwrap();
                if !(w >= prev) {
                    return Err(format!("KGraph row {} is not sorted by increasing distance (kgraph.rs:508-509)", i));
                }
                prev = w;
                csr.col.push(e.node as u32);
                csr.dist.push(w);
            }
            csr.row_ptr.push(csr.col.len() as u64);
            let id = kgraph.get_data_id_from_idx(i).ok_or_else(|| format!("no DataId for node index {}", i))?;
            csr.data_id.push(*id as u64);
        }
        Ok(csr)
    }

    pub fn write<P: AsRef<Path>>(&self, path: P) -> std::io::Result<()> {
        let mut f = BufWriter::new(File::create(path)?);
        f.write_all(MAGIC)?;
        f.write_all(&VERSION.to_le_bytes())?;
        f.write_all(&0u32.to_le_bytes())?;
        f.write_all(&(self.nb_nodes() as u64).to_le_bytes())?;
        f.write_all(&(self.nb_edges() as u64).to_le_bytes())?;
        f.write_all(&self.max_nbng.to_le_bytes())?;
        for v in &self.row_ptr {
            f.write_all(&v.to_le_bytes())?;
        }
        for v in &self.col {
            f.write_all(&v.to_le_bytes())?;
        }
        for v in &self.dist {
            f.write_all(&v.to_le_bytes())?;
        }
        for v in &self.data_id {
            f.write_all(&v.to_le_bytes())?;
        }
        f.flush()
    }

    pub fn read<P: AsRef<Path>>(path: P) -> std::io::Result<Self> {
        use std::io::{Error, ErrorKind};
        let mut f = BufReader::new(File::open(path)?);
        let mut magic = [0u8; 8];
        f.read_exact(&mut magic)?;
        if &magic != MAGIC {
            return Err(Error::new(ErrorKind::InvalidData, "not an ANNKGCSR file"));
        }
        let mut b4 = [0u8; 4];
        let mut b8 = [0u8; 8];
        f.read_exact(&mut b4)?;
        if u32::from_le_bytes(b4) != VERSION {
            return Err(Error::new(ErrorKind::InvalidData, "unsupported ANNKGCSR version"));
        }
        f.read_exact(&mut b4)?; // flags
        f.read_exact(&mut b8)?;
        let n = u64::from_le_bytes(b8) as usize;
        f.read_exact(&mut b8)?;
        let e = u64::from_le_bytes(b8) as usize;
        f.read_exact(&mut b8)?;
        let max_nbng = u64::from_le_bytes(b8);
        let mut csr = KGraphCsr { max_nbng, ..Default::default() };
        csr.row_ptr.reserve(n + 1);
        for _ in 0..=n {
            f.read_exact(&mut b8)?;
            csr.row_ptr.push(u64::from_le_bytes(b8));
        }
        csr.col.reserve(e);
        for _ in 0..e {
            f.read_exact(&mut b4)?;
            csr.col.push(u32::from_le_bytes(b4));
        }
        csr.dist.reserve(e);
        for _ in 0..e {
            f.read_exact(&mut b4)?;
            csr.dist.push(f32::from_le_bytes(b4));
        }
        csr.data_id.reserve(n);
        for _ in 0..n {
            f.read_exact(&mut b8)?;
            csr.data_id.push(u64::from_le_bytes(b8));
        }
        if csr.row_ptr.last().copied() != Some(e as u64) {
            return Err(Error::new(ErrorKind::InvalidData, "ANNKGCSR: row_ptr[n] != E"));
        }
        Ok(csr)
    }
}

/// Convenience used by the examples: dump the graph right after `kgraph_from_hnsw_all` (kgraph.rs:440-579).
pub fn write_csr<F, P: AsRef<Path>>(kgraph: &KGraph<F>, path: P) -> Result<(), String>
where
    F: FromPrimitive + Float + std::fmt::UpperExp + Sync + Send + std::iter::Sum,
{
    KGraphCsr::from_kgraph(kgraph)?.write(path).map_err(|e| e.to_string())
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn round_trip() {
        let csr = KGraphCsr {
            max_nbng: 2,
            row_ptr: vec![0, 2, 3, 5],
            col: vec![1, 2, 0, 0, 1],
            dist: vec![0.5, 1.0, 0.5, 1.0, 2.0],
            data_id: vec![10, 11, 12],
        };
        let path = std::env::temp_dir().join("annkgcsr_round_trip.bin");
        csr.write(&path).unwrap();
        assert_eq!(KGraphCsr::read(&path).unwrap(), csr);
    }
}
