// TEST INFRASTRUCTURE ONLY -- shim for vg utils.hpp.
#pragma once
#include "utility.hpp"
