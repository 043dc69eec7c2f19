// Minimal stand-in for <Rcpp.h>, TEST INFRASTRUCTURE ONLY.
//
// The reference's SPLASH.cpp / SPLASH.h include <Rcpp.h> for exactly three things
// (reference src/SPLASH.cpp:10-21, src/SPLASH.h:65,93-97):
//   * RCPP_MODULE(...) { class_<SPLASH>("SPLASH").constructor<..>().method(..)...; }
//   * Rcpp::List as the return type of spin_up / run_one_year / run_all
//   * List::create(Named("x") = std::vector<double>, ...)
// R and Rcpp are not installed in this image, so oracle/Makefile puts this header on the
// include path and compiles the reference sources unmodified from /root/reference/src.
// Nothing here is numerical; it only carries named vectors out of the reference classes.
#ifndef SPLASH_ORACLE_RCPP_SHIM_H
#define SPLASH_ORACLE_RCPP_SHIM_H

#include <cmath>
#include <string>
#include <utility>
#include <vector>

namespace Rcpp {

struct NamedVector {
    std::string name;
    std::vector<double> values;
};

struct Named {
    std::string name;
    explicit Named(const char* n) : name(n) {}
    NamedVector operator=(const std::vector<double>& v) const { return NamedVector{name, v}; }
};

class List {
public:
    std::vector<NamedVector> items;

    template <typename... Args>
    static List create(Args&&... args) {
        List out;
        (out.items.push_back(std::forward<Args>(args)), ...);
        return out;
    }

    const std::vector<double>& operator[](const char* key) const {
        for (const auto& it : items)
            if (it.name == key) return it.values;
        static const std::vector<double> empty;
        return empty;
    }
};

// class_<T>("name").constructor<...>().method("m", &T::m)  -- registration is a no-op here.
template <typename T>
class class_ {
public:
    explicit class_(const char*) {}
    template <typename... A>
    class_& constructor() { return *this; }
    template <typename F>
    class_& method(const char*, F) { return *this; }
};

}  // namespace Rcpp

#define RCPP_MODULE(name) static void splash_oracle_unused_module_##name()

#endif
