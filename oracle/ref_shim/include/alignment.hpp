// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
#include "vg/vg.pb.h"
#include "vg/io/stream.hpp"
#include "utility.hpp"
