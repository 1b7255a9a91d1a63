// TEST INFRASTRUCTURE ONLY -- shim (process set-up hooks of vg; no-ops here).
#pragma once
namespace vg {
#ifndef REF_SHIM_VG_HOOKS
#define REF_SHIM_VG_HOOKS
inline void enable_crash_handling() {}
inline void preflight_check() {}
inline void configure_memory_allocator() {}
namespace io { inline bool register_libvg_io() { return true; } }
#endif
}
