// v2 tower: backward kernels (placeholder while the forward path is validated)
#pragma once
#ifndef CDRA_EMU
#include "v2_pw.cuh"
#include "v2_dw.cuh"
namespace cdra { namespace v2 { } }
#endif
