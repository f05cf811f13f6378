// boost/functional/hash.hpp — included by the reference's dqn.hpp:11, nothing of it is used on this path.
#pragma once
#include <functional>
