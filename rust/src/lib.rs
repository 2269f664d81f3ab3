//! Source-only shim (never compiled here: no Rust toolchain in the image). Same public modules as the reference
//! crate (`src/lib.rs:1-8`), backed by libmpc_b200.so.
extern crate nalgebra as na;
pub mod ffi;
pub mod gaussian;
pub mod mppi;
pub mod ukf;
pub mod ukf2;

/// Stands in for the `dynamics` / `cost` / `fx` / `hx` fn pointers and closures of the reference: a GPU kernel
/// cannot call host code, so the models of the reference's examples are built into the library.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DeviceModel { L = 0, NL = 1, NL6 = 2, PenLin = 16, PenNl = 17, Pen6 = 18, Nl6Ukf = 19 }
