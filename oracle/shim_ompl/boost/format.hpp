// TEST SHIM: just enough of boost::format / boost::str for the file name built in createHornEnvironment.
#pragma once
#include <string>
namespace boost
{
    class format
    {
    public:
        explicit format(const char *f) : s_(f) {}
        template <typename T>
        format &operator%(const T &v)
        {
            const auto p = s_.find('%');
            if (p != std::string::npos) s_.replace(p, 2, std::to_string(v));
            return *this;
        }
        const std::string &str() const { return s_; }
    private:
        std::string s_;
    };
    inline std::string str(const format &f) { return f.str(); }
}  // namespace boost
