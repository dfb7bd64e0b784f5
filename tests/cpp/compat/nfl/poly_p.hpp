// Include-name shim: lets sources written against the reference's <nfl/poly_p.hpp> pick up the B200 drop-in header unchanged
// (tests/cpp/Makefile builds the reference's own test programs this way).
#include <nfl_b200.hpp>
