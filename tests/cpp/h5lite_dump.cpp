// Dumps every dataset of an HDF5 weights file through crcnn_b200/cpp/h5lite.hpp:
//   h5lite_dump <file.h5> <outdir>   -> one line "name rank d0 d1 ... class size" per dataset on stdout and
//                                       <outdir>/<name>.f32 (float datasets, via LoadH5::getData like CnnBuilder::getPretrained)
#include <cstdio>
#include <fstream>
#include "../../crcnn_b200/cpp/h5lite.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    try {
        crcnn_b200::LoadH5 ld;
        ld.setFileName(argv[1]);
        for (auto &kv : ld.file().datasets()) {
            std::printf("%s %zu", kv.first.c_str(), kv.second.dims.size());
            for (auto d : kv.second.dims) std::printf(" %llu", (unsigned long long)d);
            std::printf(" %d %d\n", kv.second.type_class, kv.second.type_size);
            if (kv.second.type_class != 1) continue;
            ld.setVarName(kv.first);
            std::vector<float> v = ld.getData();
            if ((int)v.size() != ld.getSize()) return 3;
            std::ofstream o(std::string(argv[2]) + "/" + kv.first + ".f32", std::ios::binary);
            o.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * 4));
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
