// TEST INFRASTRUCTURE ONLY -- shim (declarations only).
#pragma once
#include "vg/io/stream.hpp"
