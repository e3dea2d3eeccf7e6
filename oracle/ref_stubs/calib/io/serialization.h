// ORACLE — test infrastructure only.
// Empty stand-in for the reference's calib/io/serialization.h, placed first on the include path
// when oracle/_ref/libref_ransac.so is built.  The reference's common/ransac.h includes that header
// (nlohmann-json / Boost.PFR glue, neither present in this image) but uses nothing from it; with this
// stub the reference's own RANSAC template compiles unmodified from where it lies.
#pragma once
