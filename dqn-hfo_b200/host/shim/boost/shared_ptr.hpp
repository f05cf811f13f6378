// boost/shared_ptr.hpp — the reference spells Caffe's net handles boost::shared_ptr (dqn.hpp:35).
#pragma once
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}  // namespace boost
