// boost/optional.hpp — Boost is not in this image: the reference's dqn.hpp:12 / dqn_main.cpp:140 only need
// boost::optional<T> and boost::none, which std::optional / std::nullopt provide with the same semantics.
#pragma once
#include <optional>
namespace boost {
template <class T> using optional = std::optional<T>;
inline constexpr std::nullopt_t none = std::nullopt;
}  // namespace boost
