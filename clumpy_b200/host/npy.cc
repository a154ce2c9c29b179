#include "npy.hh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace npy {

static size_t find_key(const std::string& h, const char* key) {
    size_t at = h.find(key);
    if (at == std::string::npos) throw std::runtime_error(std::string("npy: header has no '") + key + "'");
    return at;
}

Array load(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("npy_load: Unable to open file " + path);
    unsigned char pre[12];
    if (fread(pre, 1, 10, fp) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
        fclose(fp);
        throw std::runtime_error("npy: not a .npy file: " + path);
    }
    size_t hlen = pre[8] | (pre[9] << 8);
    if (pre[6] >= 2) { // format 2.0+: four-byte header length
        if (fread(pre + 10, 1, 2, fp) != 2) { fclose(fp); throw std::runtime_error("npy: truncated header"); }
        hlen |= ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24);
    }
    std::string h(hlen, '\0');
    if (fread(&h[0], 1, hlen, fp) != hlen) { fclose(fp); throw std::runtime_error("npy: truncated header"); }

    Array a;
    try {
        size_t at = find_key(h, "fortran_order");
        a.fortran_order = h.find("True", at) != std::string::npos && h.find("True", at) < h.find(',', at);
        at = find_key(h, "descr");
        at = h.find('\'', h.find(':', at)); // opening quote of the dtype string
        const char endian = h[at + 1];
        if (endian != '<' && endian != '|') throw std::runtime_error("parse_npy_header: requires little endian");
        a.type_code = h[at + 2];
        a.word_size = (size_t)atoi(h.c_str() + at + 3);
        at = find_key(h, "shape");
        size_t open = h.find('(', at), close = h.find(')', at);
        if (open == std::string::npos || close == std::string::npos) throw std::runtime_error("npy: bad shape");
        const char* p = h.c_str() + open + 1;
        const char* end = h.c_str() + close;
        while (p < end) {
            while (p < end && (*p < '0' || *p > '9')) ++p;
            if (p >= end) break;
            char* next = nullptr;
            a.shape.push_back((size_t)strtoull(p, &next, 10));
            p = next;
        }
    } catch (...) {
        fclose(fp);
        throw;
    }
    a.bytes.resize(a.num_vals() * a.word_size);
    const size_t got = a.bytes.empty() ? 0 : fread(a.bytes.data(), 1, a.bytes.size(), fp);
    fclose(fp);
    if (got != a.bytes.size()) throw std::runtime_error("load_the_npy_file: failed fread");
    return a;
}

std::vector<char> make_header(char type_code, size_t word_size, const std::vector<size_t>& shape) {
    std::string dict = "{'descr': '<";
    dict += type_code;
    dict += std::to_string(word_size);
    dict += "', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < shape.size(); ++i) {
        if (i) dict += ", ";
        dict += std::to_string(shape[i]);
    }
    if (shape.size() == 1) dict += ",";
    dict += "), }";
    // cnpy pads with 16 - (10 + len) % 16 spaces (a full 16 when already aligned); last byte '\n'
    const size_t pad = 16 - (10 + dict.size()) % 16;
    dict.append(pad, ' ');
    dict.back() = '\n';
    std::vector<char> out = {(char)0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (char)(dict.size() & 0xFF), (char)(dict.size() >> 8)};
    out.insert(out.end(), dict.begin(), dict.end());
    return out;
}

bool save(const std::string& path, char type_code, size_t word_size, const void* data,
          const std::vector<size_t>& shape) {
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    const std::vector<char> header = make_header(type_code, word_size, shape);
    size_t n = word_size;
    for (size_t s : shape) n *= s;
    bool ok = fwrite(header.data(), 1, header.size(), fp) == header.size();
    ok = ok && (n == 0 || fwrite(data, 1, n, fp) == n);
    return (fclose(fp) == 0) && ok;
}

} // namespace npy
