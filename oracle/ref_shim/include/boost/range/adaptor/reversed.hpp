// TEST INFRASTRUCTURE ONLY -- shim (the reference includes it, uses nothing of it on this path).
#pragma once
