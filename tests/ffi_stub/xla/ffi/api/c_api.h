// TEST INFRASTRUCTURE ONLY - stand-in for XLA's xla/ffi/api/c_api.h (see ffi.h next to this file).
#pragma once
typedef struct XLA_FFI_CallFrame XLA_FFI_CallFrame;
typedef struct XLA_FFI_Error XLA_FFI_Error;
