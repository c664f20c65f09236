// boost/serialization/access.hpp -- stand-in: the friend declaration target only.
#pragma once
namespace boost { namespace serialization { class access {}; } }
