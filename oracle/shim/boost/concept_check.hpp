#pragma once
#define BOOST_CONCEPT_USAGE(name) void name##_concept_usage()
#define BOOST_CONCEPT_ASSERT(x) static_assert(true, "")
