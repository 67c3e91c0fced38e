// ORACLE — TEST INFRASTRUCTURE ONLY.  fcl::Triangle index triple (src/planning/fcl_utils.cpp:27-34).
#pragma once
#include <cstddef>
namespace fcl {
struct Triangle {
	std::size_t i[3]{0, 0, 0};
	Triangle() = default;
	Triangle(std::size_t a, std::size_t b, std::size_t c) : i{a, b, c} {}
};
}  // namespace fcl
