// TEST INFRASTRUCTURE ONLY -- shim: opaque gbwt types held by pointer in Trailmix_struct.h (unused on this path).
#pragma once
namespace gbwt { class GBWT {}; }
namespace gbwtgraph { class GBWTGraph {}; }
