// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
#include "bdsg/odgi.hpp"
