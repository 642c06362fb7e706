// Stand-in: DBoW2's FeatureVector.h / BowVector.h only befriend the archive access class and name base_object inside
// a serialize() template that is never instantiated here.
#ifndef ORBREF_STUB_BOOST_SERIALIZATION_HPP_
#define ORBREF_STUB_BOOST_SERIALIZATION_HPP_
namespace boost {
namespace serialization {
class access;
template <class B, class D>
B& base_object(D& d) { return static_cast<B&>(d); }
}  // namespace serialization
}  // namespace boost
#endif
