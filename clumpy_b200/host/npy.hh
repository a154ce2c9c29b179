// npy.hh -- minimal .npy (format 1.0 / 2.0) reader and writer, byte-compatible with what the
// reference reads and writes through its vendored cnpy (extern/cnpy/cnpy.h:89-133,237-266 and
// cnpy.cpp:104-153): little-endian, C order, header padded to a multiple of 16 bytes.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace npy {

struct Array {
    std::vector<size_t> shape;
    size_t word_size = 0;   // bytes per element
    char type_code = '?';   // 'f', 'u', 'i', ...
    bool fortran_order = false;
    std::vector<char> bytes;

    template <typename T> const T* data() const { return reinterpret_cast<const T*>(bytes.data()); }
    size_t num_vals() const {
        size_t n = 1;
        for (size_t s : shape) n *= s;
        return n;
    }
};

// Throws std::runtime_error on a missing / malformed file, like cnpy::npy_load.
Array load(const std::string& path);

// The exact header bytes cnpy::create_npy_header produces for this dtype and shape.
std::vector<char> make_header(char type_code, size_t word_size, const std::vector<size_t>& shape);

// Writes header + raw data ("w" mode of cnpy::npy_save).  Returns false if the file cannot be written.
bool save(const std::string& path, char type_code, size_t word_size, const void* data,
          const std::vector<size_t>& shape);

inline bool save_f32(const std::string& p, const float* d, const std::vector<size_t>& s) { return save(p, 'f', 4, d, s); }
inline bool save_u8(const std::string& p, const uint8_t* d, const std::vector<size_t>& s) { return save(p, 'u', 1, d, s); }

} // namespace npy
