// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
namespace vg { namespace subcommand { class Subcommand { public: template <class... A> Subcommand(A...) {} }; } }
