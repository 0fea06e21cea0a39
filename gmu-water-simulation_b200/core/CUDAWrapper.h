// CUDAWrapper.h — C++ convenience layer over the C ABI (include/sph_cuda.h).
//
// Plays the role CLWrapper/CLPlatforms play in the reference (include/CLWrapper.h:24-69,
// include/CLPlatforms.h:7-18): owns the device context, turns status codes into exceptions
// (CUDAException ≙ CLException, include/CLWrapper.h:18-22) and lists devices for the GUI combo box.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sph_cuda.h"

class CUDAException : public std::runtime_error {
public:
    explicit CUDAException(const std::string &what) : std::runtime_error("CUDA::" + what) {}
};

class CUDAPlatforms {
public:
    static int getDeviceCount() {
        int n = 0;
        if (sph_device_count(&n) != SPH_OK) return 0;
        return n;
    }
    static std::string getDeviceInfo(int device) {
        char buf[256];
        if (sph_device_name(device, buf, sizeof(buf)) != SPH_OK) throw CUDAException(sph_last_error(nullptr));
        return buf;
    }
    static std::vector<std::string> getDevices() {
        std::vector<std::string> out;
        for (int d = 0; d < getDeviceCount(); ++d) out.push_back(getDeviceInfo(d));
        return out;
    }
};

class CUDAWrapper {
public:
    explicit CUDAWrapper(const sph_config &cfg, bool slab = false) {
        const int rc = slab ? sph_slab_create(&cfg, &m_ctx) : sph_create(&cfg, &m_ctx);
        if (rc != SPH_OK) throw CUDAException(sph_last_error(nullptr));
    }
    ~CUDAWrapper() { sph_destroy(m_ctx); }
    CUDAWrapper(const CUDAWrapper &) = delete;
    CUDAWrapper &operator=(const CUDAWrapper &) = delete;

    sph_context *ctx() const { return m_ctx; }
    // ≙ CLWrapper::checkError (src/CLWrapper.cpp:115-121)
    void check(int status, const char *where) const {
        if (status != SPH_OK) throw CUDAException(std::string(sph_last_error(m_ctx)) + " | " + where);
    }

private:
    sph_context *m_ctx = nullptr;
};
