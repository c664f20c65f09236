// boost/serialization/split_member.hpp -- stand-in: serialisation is never executed by the pinned functions.
#pragma once
#include "access.hpp"
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
