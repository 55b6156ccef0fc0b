// host_internal.h — what api.cu needs from host.cpp (not exported).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/asgart_b200.h"

// families as the C ABI hands them out (asgart_b200_result_*): CSR offsets + ProtoSD rows
struct asgart_b200_result {
    std::vector<uint64_t> fam_off;
    std::vector<asgart_b200_protosd> sds;
};

// prepare_data result whose strand lives only on the device (GPU-side FASTA ingest): fragment map + chunks + length.
asgart_b200_prepared* ab200_prepared_device_only(const std::string& file_names, uint64_t n, const std::vector<std::string>& names,
                                                 const std::vector<uint64_t>& pos, const std::vector<uint64_t>& len,
                                                 const std::vector<asgart_b200_chunk>& chunks);
