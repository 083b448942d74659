#include <string>
namespace bb { void fault_point() {} void set_tls_error(const std::string&) {} }
