// oracle/refstub — TEST INFRASTRUCTURE: the two boost::serialization names Thirdparty/DBoW2/DBoW2/{BowVector,FeatureVector}.h mention
// (inside templates that are never instantiated here), so that the reference's own DBoW2 headers compile without boost.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base &base_object(Derived &d) { return static_cast<Base &>(d); }
}}
