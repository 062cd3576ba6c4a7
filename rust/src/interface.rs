//! The reference's public API (mcomstock/the-tessellator src/interface.rs) over the CUDA library.
//!
//! Same type and method names and the same argument meaning as the reference: `Diagram` (:25),
//! `Cell` (:237), `VoronoiFace` (:393).  Differences, all forced by defects of the reference
//!: `add_particle_with_group` / `initialize` are `pub` (D3); a cell's start
//! `Polyhedron` must be the diagram's container box; surviving container walls are reported as
//! neighbours `usize::MAX - 5 ..= usize::MAX` (F, R, B, L, U, D) instead of panicking (D10).
//! The first `compute_voronoi_cell` for a given `(search_radius, target_group)` computes every
//! cell of the diagram on the GPU in one batch; later cells read their row.
use crate::ffi;
use std::cell::RefCell;
use std::ffi::CStr;
use std::ptr;
use std::rc::Rc;

/// celery.rs:56-60
pub trait ToCeleryPoint<FloatType> {
    fn get_x(&self) -> FloatType;
    fn get_y(&self) -> FloatType;
    fn get_z(&self) -> FloatType;
}

/// vector3.rs:24-28 (f64 only: the reference implements only Float64, float.rs:78)
#[derive(Clone, Copy, Debug, Default, PartialEq)]
#[repr(C)]
pub struct Vector3 {
    pub x: f64,
    pub y: f64,
    pub z: f64,
}

/// Polyhedron::new (polyhedron.rs:226-233): the axis-aligned start box.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Polyhedron {
    pub x_min: f64,
    pub y_min: f64,
    pub z_min: f64,
    pub x_max: f64,
    pub y_max: f64,
    pub z_max: f64,
}
impl Polyhedron {
    pub fn new(x_min: f64, y_min: f64, z_min: f64, x_max: f64, y_max: f64, z_max: f64) -> Polyhedron {
        Polyhedron { x_min, y_min, z_min, x_max, y_max, z_max }
    }
}

fn check(rc: i32) {
    if rc != ffi::TESS_OK {
        // the reference panics on every error (unwrap / get_or_fail); so does this layer
        let msg = unsafe { CStr::from_ptr(ffi::tess_last_error()) }.to_string_lossy().into_owned();
        panic!("libtess_b200 error {}: {}", rc, msg);
    }
}

struct Batch {
    r: *mut ffi::tess_result,
    has_vertices: bool,
    volumes: *const f64,
    face_offsets: *const u64,
    neighbors: *const i64,
    areas: *const f64,
    // TESS_OUT_VERTICES: per-cell vertex lists and per-face loops (indices into the owning cell's list)
    vertex_offsets: *const u64,
    vertices: *const f64,
    face_vertex_offsets: *const u64,
    face_vertex_indices: *const u32,
}
impl Drop for Batch {
    fn drop(&mut self) {
        unsafe { ffi::tess_result_free(self.r) }
    }
}

/// interface.rs:25
pub struct Diagram<PointType: ToCeleryPoint<f64>> {
    d: *mut ffi::tess_diagram,
    initialized: bool,
    points: Vec<Vector3>,
    groups: Vec<u64>,
    container: Option<Polyhedron>,
    batches: RefCell<Vec<(Option<u64>, Option<usize>, Rc<Batch>)>>,
    _marker: std::marker::PhantomData<PointType>,
}

impl<PointType: ToCeleryPoint<f64>> Default for Diagram<PointType> {
    fn default() -> Self {
        let mut d = ptr::null_mut();
        check(unsafe { ffi::tess_diagram_create(&mut d, ffi::TESS_F64, 0) });
        Diagram { d, initialized: false, points: Vec::new(), groups: Vec::new(), container: None, batches: RefCell::new(Vec::new()), _marker: std::marker::PhantomData }
    }
}
impl<PointType: ToCeleryPoint<f64>> Drop for Diagram<PointType> {
    fn drop(&mut self) {
        self.batches.borrow_mut().clear();
        unsafe { ffi::tess_diagram_destroy(self.d) }
    }
}

impl<PointType: ToCeleryPoint<f64>> Diagram<PointType> {
    /// interface.rs:52-57
    pub fn add_particle_with_group(&mut self, particle: PointType, group: usize) {
        debug_assert!(!self.initialized);
        self.points.push(Vector3 { x: particle.get_x(), y: particle.get_y(), z: particle.get_z() });
        self.groups.push(group as u64);
    }
    /// Container given explicitly (the reference's `container_shape`, interface.rs:30).
    pub fn set_container(&mut self, container: Polyhedron) {
        self.container = Some(container);
    }
    /// interface.rs:60-84
    pub fn initialize(&mut self) {
        debug_assert!(!self.initialized);
        check(unsafe {
            ffi::tess_diagram_add_particles(self.d, self.points.as_ptr() as *const _, self.points.len(), std::mem::size_of::<Vector3>(), self.groups.as_ptr(), ptr::null_mut())
        });
        match self.container {
            Some(c) => {
                let b = [c.x_min, c.y_min, c.z_min, c.x_max, c.y_max, c.z_max];
                check(unsafe { ffi::tess_diagram_initialize(self.d, b.as_ptr(), ptr::null_mut()) });
            }
            None => {
                check(unsafe { ffi::tess_diagram_initialize(self.d, ptr::null(), ptr::null_mut()) });
                let mut b = [0f64; 6];
                check(unsafe { ffi::tess_diagram_grid_info(self.d, ptr::null_mut(), ptr::null_mut(), b.as_mut_ptr(), ptr::null_mut(), ptr::null_mut()) });
                self.container = Some(Polyhedron::new(b[0], b[2], b[4], b[1], b[3], b[5]));
            }
        }
        self.initialized = true;
    }

    fn opts(search_radius: Option<f64>, target_group: Option<usize>, vertices: bool) -> ffi::tess_opts {
        let mut o = std::mem::MaybeUninit::<ffi::tess_opts>::uninit();
        let mut o = unsafe {
            ffi::tess_opts_default(o.as_mut_ptr());
            o.assume_init()
        };
        if let Some(r) = search_radius {
            o.search_radius = r;
        }
        if let Some(g) = target_group {
            o.target_group = g as i64;
        }
        // the geometry outputs (vertex lists, face loops) are computed only for Cell::compute_vertices /
        // VoronoiFace::compute_vertices: the first such call re-computes the batch with them
        if vertices {
            o.outputs |= ffi::TESS_OUT_VERTICES;
        }
        o
    }
    fn wrap(r: *mut ffi::tess_result, vertices: bool) -> Rc<Batch> {
        let mut b = Batch {
            r,
            has_vertices: vertices,
            volumes: ptr::null(),
            face_offsets: ptr::null(),
            neighbors: ptr::null(),
            areas: ptr::null(),
            vertex_offsets: ptr::null(),
            vertices: ptr::null(),
            face_vertex_offsets: ptr::null(),
            face_vertex_indices: ptr::null(),
        };
        unsafe {
            check(ffi::tess_result_volumes(r, &mut b.volumes));
            check(ffi::tess_result_face_offsets(r, &mut b.face_offsets));
            check(ffi::tess_result_neighbors(r, &mut b.neighbors));
            check(ffi::tess_result_areas(r, &mut b.areas));
            if vertices {
                check(ffi::tess_result_vertex_offsets(r, &mut b.vertex_offsets));
                check(ffi::tess_result_vertices(r, &mut b.vertices));
                check(ffi::tess_result_face_vertex_offsets(r, &mut b.face_vertex_offsets));
                check(ffi::tess_result_face_vertex_indices(r, &mut b.face_vertex_indices));
            }
        }
        Rc::new(b)
    }
    fn batch(&self, search_radius: Option<f64>, target_group: Option<usize>, vertices: bool) -> Rc<Batch> {
        let key = search_radius.map(|r| r.to_bits());
        if let Some(hit) = self.batches.borrow().iter().find(|(r, g, b)| *r == key && *g == target_group && (b.has_vertices || !vertices)) {
            return hit.2.clone();
        }
        let o = Self::opts(search_radius, target_group, vertices);
        let mut r = ptr::null_mut();
        check(unsafe { ffi::tess_compute_all(self.d, &o, &mut r) });
        let b = Self::wrap(r, vertices);
        // a batch with geometry replaces the one without (cells that hold the old one keep it alive)
        self.batches.borrow_mut().retain(|(r, g, _)| !(*r == key && *g == target_group));
        self.batches.borrow_mut().push((key, target_group, b.clone()));
        b
    }

    /// Extension: every cell of the diagram in one batch (what the first `compute_voronoi_cell` triggers anyway).
    pub fn compute_all_cells(&self, search_radius: Option<f64>, target_group: Option<usize>) {
        let _ = self.batch(search_radius, target_group, false);
    }
    /// Number of particles added so far.
    pub fn len(&self) -> usize {
        self.points.len()
    }

    /// interface.rs:186-208
    pub fn get_cell_at_index(&self, index: usize, polyhedron: Polyhedron, search_radius: Option<f64>, target_group: Option<usize>) -> Cell<PointType> {
        assert_eq!(Some(polyhedron), self.container, "the start polyhedron of a cell must be the diagram's container box");
        Cell { diagram: self, index: Some(index), position: self.points[index], search_radius, target_group, batch: None, row: index }
    }
    /// interface.rs:211-232
    pub fn get_cell_at_particle(&self, point: PointType, polyhedron: Polyhedron, search_radius: Option<f64>, target_group: Option<usize>) -> Cell<PointType> {
        assert_eq!(Some(polyhedron), self.container, "the start polyhedron of a cell must be the diagram's container box");
        let position = Vector3 { x: point.get_x(), y: point.get_y(), z: point.get_z() };
        Cell { diagram: self, index: None, position, search_radius, target_group, batch: None, row: 0 }
    }
}

/// interface.rs:237
pub struct Cell<'a, PointType: ToCeleryPoint<f64>> {
    diagram: &'a Diagram<PointType>,
    index: Option<usize>,
    position: Vector3,
    search_radius: Option<f64>,
    target_group: Option<usize>,
    batch: Option<Rc<Batch>>,
    row: usize,
}

/// Neighbour id of a surviving container wall: `usize::MAX - 5 ..= usize::MAX` for F, R, B, L, U, D.
pub fn wall_neighbor(id: i64) -> usize {
    if id >= 0 { id as usize } else { usize::max_value() - 6 + (-id) as usize }
}

impl<'a, PointType: ToCeleryPoint<f64>> Cell<'a, PointType> {
    /// interface.rs:257-313
    pub fn compute_voronoi_cell(&mut self) {
        self.compute(false);
    }
    fn compute(&mut self, vertices: bool) {
        match self.index {
            Some(i) => {
                self.batch = Some(self.diagram.batch(self.search_radius, self.target_group, vertices));
                self.row = i;
            }
            None => {
                let o = Diagram::<PointType>::opts(self.search_radius, self.target_group, vertices);
                let mut r = ptr::null_mut();
                check(unsafe { ffi::tess_compute_at_points(self.diagram.d, &self.position as *const Vector3 as *const f64, 1, &o, &mut r) });
                self.batch = Some(Diagram::<PointType>::wrap(r, vertices));
                self.row = 0;
            }
        }
    }
    fn need(&mut self) -> Rc<Batch> {
        if self.batch.is_none() {
            self.compute(false);
        }
        self.batch.as_ref().unwrap().clone()
    }
    /// the batch WITH the geometry outputs: computed on the first call that reads vertices
    fn need_vertices(&mut self) -> Rc<Batch> {
        let have = match self.batch.as_ref() {
            Some(b) => b.has_vertices,
            None => false,
        };
        if !have {
            self.compute(true);
        }
        self.batch.as_ref().unwrap().clone()
    }
    fn range(b: &Batch, row: usize) -> (usize, usize) {
        unsafe { (*b.face_offsets.add(row) as usize, *b.face_offsets.add(row + 1) as usize) }
    }
    /// interface.rs:337-339
    pub fn compute_volume(&mut self) -> f64 {
        let b = self.need();
        unsafe { *b.volumes.add(self.row) }
    }
    /// interface.rs:342-344
    pub fn compute_neighbors(&mut self) -> Vec<usize> {
        let b = self.need();
        let (lo, hi) = Self::range(&b, self.row);
        (lo..hi).map(|k| wall_neighbor(unsafe { *b.neighbors.add(k) })).collect()
    }
    /// interface.rs:348-365: `ExpandingSearch::expand_all_in_radius(radius)` around the cell's position (particle
    /// indices in search-table order), then only the members of `target_group` if one is given.
    pub fn compute_neighbor_cloud(&self, radius: f64, target_group: Option<usize>) -> Vec<usize> {
        let mut q = ptr::null_mut();
        let group = match target_group {
            Some(g) => g as i64,
            None => -1,
        };
        check(unsafe {
            ffi::tess_find_neighbors(self.diagram.d, &self.position as *const Vector3 as *const f64, 1, radius, ffi::TESS_QUERY_NEIGHBOR_CLOUD, group, ptr::null_mut(), &mut q)
        });
        let mut offsets: *const u64 = ptr::null();
        let mut indices: *const i64 = ptr::null();
        let (rc1, rc2) = unsafe { (ffi::tess_query_offsets(q, &mut offsets), ffi::tess_query_indices(q, &mut indices)) };
        let mut out = Vec::new();
        if rc1 == ffi::TESS_OK && rc2 == ffi::TESS_OK {
            let (lo, hi) = unsafe { (*offsets as usize, *offsets.add(1) as usize) };
            out.extend((lo..hi).map(|k| (unsafe { *indices.add(k) }) as usize));
        }
        unsafe { ffi::tess_query_free(q) };
        check(rc1);
        check(rc2);
        out
    }
    /// interface.rs:368-370: the vertices of the cell, in cell-local coordinates (relative to the particle, as the
    /// reference's polyhedron is translated by -position, interface.rs:266).
    pub fn compute_vertices(&mut self) -> Vec<Vector3> {
        let b = self.need_vertices();
        let (lo, hi) = unsafe { (*b.vertex_offsets.add(self.row) as usize, *b.vertex_offsets.add(self.row + 1) as usize) };
        (lo..hi).map(|v| unsafe { Vector3 { x: *b.vertices.add(3 * v), y: *b.vertices.add(3 * v + 1), z: *b.vertices.add(3 * v + 2) } }).collect()
    }
    /// interface.rs:373-384.  Faces carry the geometry batch (VoronoiFace::compute_vertices reads the loops): asking for
    /// the faces is the request for geometry; volumes, neighbours and areas alone never compute it.
    pub fn compute_faces(&mut self) -> Vec<VoronoiFace> {
        let b = self.need_vertices();
        let (lo, hi) = Self::range(&b, self.row);
        let row = self.row;
        (lo..hi).map(|k| VoronoiFace { batch: b.clone(), k, row }).collect()
    }
    /// interface.rs:387-389
    pub fn original_index(&self) -> Option<usize> {
        self.index
    }
}

/// interface.rs:393
pub struct VoronoiFace {
    batch: Rc<Batch>,
    k: usize,
    row: usize,
}
impl VoronoiFace {
    /// interface.rs:403-405 -> Polyhedron::compute_face_vertices (polyhedron.rs:897-919): the face's vertices in loop
    /// order, starting at the target of the face's starting edge.
    pub fn compute_vertices(&self) -> Vec<Vector3> {
        let b = &self.batch;
        let base = unsafe { *b.vertex_offsets.add(self.row) as usize };
        let (lo, hi) = unsafe { (*b.face_vertex_offsets.add(self.k) as usize, *b.face_vertex_offsets.add(self.k + 1) as usize) };
        (lo..hi)
            .map(|i| unsafe {
                let v = base + *b.face_vertex_indices.add(i) as usize;
                Vector3 { x: *b.vertices.add(3 * v), y: *b.vertices.add(3 * v + 1), z: *b.vertices.add(3 * v + 2) }
            })
            .collect()
    }
    /// interface.rs:408-410
    pub fn compute_area(&self) -> f64 {
        unsafe { *self.batch.areas.add(self.k) }
    }
    /// interface.rs:413-416
    pub fn compute_neighbor(&self) -> usize {
        wall_neighbor(unsafe { *self.batch.neighbors.add(self.k) })
    }
}
