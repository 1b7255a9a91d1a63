// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
#include "gbwtgraph/gbwtgraph.h"
