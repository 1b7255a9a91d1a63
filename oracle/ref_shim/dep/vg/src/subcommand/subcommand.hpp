// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
#include "subcommand/subcommand.hpp"
