// Forward declarations only (see static_vector.hpp in ../container).
#pragma once
namespace boost::stacktrace {
class frame;
class stacktrace;
}  // namespace boost::stacktrace
