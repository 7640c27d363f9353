// oracle/ref_strset_shim.cpp — TEST INFRASTRUCTURE ONLY.
// glue only: exposes the reference's ordered_set<> over StringList64 (src/hash_string.hpp, src/superstring.hpp, unmodified, included
// from where they lie) to Python through plain numpy buffers, because the reference's own binding of StringList64 lives in the
// superstrings module, which needs pcre and is not built here.
#include "hash_string.hpp"
namespace py = pybind11;
using namespace vaex;

static std::shared_ptr<StringList64> make_list(py::array_t<int64_t> offsets, py::array_t<uint8_t> bytes, py::object mask) {
    const int64_t n = offsets.shape(0) - 1;
    const int64_t nbytes = offsets.at(n);
    auto sl = std::make_shared<StringList64>(nbytes, n);
    std::copy(bytes.data(), bytes.data() + nbytes, (uint8_t *)sl->bytes);
    for (int64_t i = 0; i <= n; i++)
        sl->indices[i] = offsets.at(i);
    if (!mask.is_none()) {
        py::array_t<uint8_t> m = mask.cast<py::array_t<uint8_t>>();
        sl->ensure_null_bitmap();
        for (int64_t i = 0; i < n; i++)
            if (m.at(i))
                sl->set_null(i);
    }
    return sl;
}

PYBIND11_MODULE(strset_shim, m) {
    py::class_<ordered_set<>>(m, "ordered_set_string")
        .def(py::init<int, int64_t>(), py::arg("nmaps"), py::arg("limit") = -1)
        .def("update", [](ordered_set<> &s, py::array_t<int64_t> offsets, py::array_t<uint8_t> bytes, py::object mask, int64_t start_index, bool return_values) {
            auto sl = make_list(offsets, bytes, mask);
            return s.update(sl.get(), start_index, 1024 * 128, 1024 * 128, return_values);
        })
        .def("map_ordinal", [](ordered_set<> &s, py::array_t<int64_t> offsets, py::array_t<uint8_t> bytes, py::object mask) {
            auto sl = make_list(offsets, bytes, mask);
            return s.map_ordinal(sl.get());
        })
        .def("key_array", [](ordered_set<> &s) {
            auto sl = s.key_array();
            const int64_t n = sl->length;
            py::array_t<int64_t> off(n + 1);
            for (int64_t i = 0; i <= n; i++)
                off.mutable_at(i) = sl->indices[i] - sl->indices[0];
            const int64_t nb = sl->indices[n] - sl->indices[0];
            py::array_t<uint8_t> by(nb);
            std::copy(sl->bytes + sl->indices[0], sl->bytes + sl->indices[0] + nb, (char *)by.mutable_data());
            py::array_t<uint8_t> nulls(n);
            for (int64_t i = 0; i < n; i++)
                nulls.mutable_at(i) = sl->is_null(i);
            return py::make_tuple(off, by, nulls);
        })
        .def("offsets", &ordered_set<>::offsets)
        .def("__len__", &ordered_set<>::length)
        .def_property_readonly("null_count", [](const ordered_set<> &c) { return c.null_count; })
        .def_property_readonly("null_index", [](const ordered_set<> &c) { return c.null_index(); });
    m.def("hash", [](py::bytes b) { std::string s = b; return (uint64_t)std::hash<string_view>()(string_view(s)); });
}
