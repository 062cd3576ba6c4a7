//! the-tessellator-b200: the reference crate's `interface` module backed by the B200 CUDA library.
pub mod ffi;
pub mod interface;
