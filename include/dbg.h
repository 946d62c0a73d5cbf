// dbg.h -- minimal stand-in for the vendored sharkdp/dbg-macro of the reference (include/dbg.h,
// third-party there).  Keeps the one-line stderr format the reference's run scripts scrape with
// grep/awk (Figure10/run.sh:17-22):   [file:line (function)] expression = value (type)
#ifndef GNNAGG_COMPAT_DBG_H
#define GNNAGG_COMPAT_DBG_H
#include <cstring>
#include <iostream>
#include <string>
#include <typeinfo>

namespace gnnagg_dbg {
template <class T>
struct type_label { static const char *get() { return typeid(T).name(); } };
#define GNNAGG_DBG_LABEL(T, S) template <> struct type_label<T> { static const char *get() { return S; } };
GNNAGG_DBG_LABEL(int, "int")
GNNAGG_DBG_LABEL(unsigned, "unsigned int")
GNNAGG_DBG_LABEL(long, "long")
GNNAGG_DBG_LABEL(unsigned long, "unsigned long")
GNNAGG_DBG_LABEL(float, "float")
GNNAGG_DBG_LABEL(double, "double")
GNNAGG_DBG_LABEL(bool, "bool")
GNNAGG_DBG_LABEL(std::string, "std::string")
#undef GNNAGG_DBG_LABEL

template <class T>
inline T &&show(const char *file, int line, const char *func, const char *expr, T &&value)
{
    const char *base = std::strrchr(file, '/');
    std::cerr << '[' << (base ? base + 1 : file) << ':' << line << " (" << func << ")] " << expr << " = " << value
              << " (" << type_label<typename std::decay<T>::type>::get() << ")\n";
    return static_cast<T &&>(value);
}
template <size_t N>
inline const char *show(const char *file, int line, const char *func, const char *, const char (&literal)[N])
{
    const char *base = std::strrchr(file, '/');
    std::cerr << '[' << (base ? base + 1 : file) << ':' << line << " (" << func << ")] " << literal << '\n';
    return literal;
}
}  // namespace gnnagg_dbg

#define dbg(...) gnnagg_dbg::show(__FILE__, __LINE__, __func__, #__VA_ARGS__, (__VA_ARGS__))
#endif
