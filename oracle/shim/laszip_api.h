/* oracle/shim/laszip_api.h — TEST INFRASTRUCTURE ONLY.
 *
 * In-memory stand-in (our own code) for the LASzip C API that the reference's LAS reader / writer
 * (core/io/LASFile.*, core/io/LASPersistence.*) are written against.  LASzip is a third-party dependency
 * that is neither vendored in /root/reference nor installed in this image; this header declares the
 * subset of its public API those files use, so that oracle/Makefile can compile them VERBATIM and the
 * oracle can observe what LASPersistence::persist_points hands to the LAS writer:
 *   - "files" are records in a process-wide map keyed by path (header + written points);
 *   - laszip_set_coordinates restates LASzip's published quantisation
 *       X = I32_QUANTIZE((coordinate - offset) / scale_factor),
 *       I32_QUANTIZE(n) = (n >= 0) ? (I32)(n + 0.5) : (I32)(n - 0.5)      (LASzip mydefs.hpp, laszip_dll.cpp)
 *     which is the one piece of this path whose parity is anchored on LASzip's documentation instead of
 *     on code compiled from the reference tree.
 * No compression, no file IO.  Nothing here is used by the product.
 */
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

typedef int laszip_BOOL;
typedef unsigned char laszip_U8;
typedef unsigned short laszip_U16;
typedef unsigned int laszip_U32;
typedef unsigned long long laszip_U64;
typedef signed char laszip_I8;
typedef short laszip_I16;
typedef int laszip_I32;
typedef long long laszip_I64;
typedef char laszip_CHAR;
typedef float laszip_F32;
typedef double laszip_F64;
typedef void* laszip_POINTER;

struct laszip_header
{
  laszip_U16 file_source_ID = 0;
  laszip_U16 global_encoding = 0;
  laszip_U32 project_ID_GUID_data_1 = 0;
  laszip_U16 project_ID_GUID_data_2 = 0;
  laszip_U16 project_ID_GUID_data_3 = 0;
  laszip_CHAR project_ID_GUID_data_4[8] = {};
  laszip_U8 version_major = 1;
  laszip_U8 version_minor = 2;
  laszip_CHAR system_identifier[32] = {};
  laszip_CHAR generating_software[32] = {};
  laszip_U16 file_creation_day = 0;
  laszip_U16 file_creation_year = 0;
  laszip_U16 header_size = 227;
  laszip_U32 offset_to_point_data = 227;
  laszip_U32 number_of_variable_length_records = 0;
  laszip_U8 point_data_format = 0;
  laszip_U16 point_data_record_length = 20;
  laszip_U32 number_of_point_records = 0;
  laszip_U32 number_of_points_by_return[5] = {};
  laszip_F64 x_scale_factor = 0.01;
  laszip_F64 y_scale_factor = 0.01;
  laszip_F64 z_scale_factor = 0.01;
  laszip_F64 x_offset = 0;
  laszip_F64 y_offset = 0;
  laszip_F64 z_offset = 0;
  laszip_F64 max_x = 0;
  laszip_F64 min_x = 0;
  laszip_F64 max_y = 0;
  laszip_F64 min_y = 0;
  laszip_F64 max_z = 0;
  laszip_F64 min_z = 0;
  laszip_U64 extended_number_of_point_records = 0;
  laszip_U64 extended_number_of_points_by_return[15] = {};
};

struct laszip_point
{
  laszip_I32 X = 0;
  laszip_I32 Y = 0;
  laszip_I32 Z = 0;
  laszip_U16 intensity = 0;
  laszip_U8 return_number : 3;
  laszip_U8 number_of_returns : 3;
  laszip_U8 scan_direction_flag : 1;
  laszip_U8 edge_of_flight_line : 1;
  laszip_U8 classification : 5;
  laszip_U8 synthetic_flag : 1;
  laszip_U8 keypoint_flag : 1;
  laszip_U8 withheld_flag : 1;
  laszip_I8 scan_angle_rank = 0;
  laszip_U8 user_data = 0;
  laszip_U16 point_source_ID = 0;
  laszip_F64 gps_time = 0;
  laszip_U16 rgb[4] = {};
  laszip_point()
    : return_number(0)
    , number_of_returns(0)
    , scan_direction_flag(0)
    , edge_of_flight_line(0)
    , classification(0)
    , synthetic_flag(0)
    , keypoint_flag(0)
    , withheld_flag(0)
  {}
};

namespace laszip_shim {
struct File
{
  laszip_header header;
  std::vector<laszip_point> points;
};
struct Handle
{
  laszip_header header;
  laszip_point point;
  File* file = nullptr;
  size_t cursor = 0;
  bool writing = false;
  std::string error;
};
inline std::map<std::string, File>&
files()
{
  static std::map<std::string, File> f;
  return f;
}
inline laszip_I32
quantize(double n)
{
  return (n >= 0) ? static_cast<laszip_I32>(n + 0.5) : static_cast<laszip_I32>(n - 0.5);
}
} // namespace laszip_shim

inline laszip_I32
laszip_create(laszip_POINTER* pointer)
{
  *pointer = new laszip_shim::Handle();
  return 0;
}
inline laszip_I32
laszip_destroy(laszip_POINTER pointer)
{
  delete static_cast<laszip_shim::Handle*>(pointer);
  return 0;
}
inline laszip_I32
laszip_get_error(laszip_POINTER pointer, laszip_CHAR** error)
{
  static char none[] = "laszip shim: no error";
  *error = none;
  return 0;
}
inline laszip_I32
laszip_get_header_pointer(laszip_POINTER pointer, laszip_header** header_pointer)
{
  *header_pointer = &static_cast<laszip_shim::Handle*>(pointer)->header;
  return 0;
}
inline laszip_I32
laszip_get_point_pointer(laszip_POINTER pointer, laszip_point** point_pointer)
{
  *point_pointer = &static_cast<laszip_shim::Handle*>(pointer)->point;
  return 0;
}
inline laszip_I32
laszip_set_header(laszip_POINTER pointer, const laszip_header* header)
{
  static_cast<laszip_shim::Handle*>(pointer)->header = *header;
  return 0;
}
inline laszip_I32
laszip_set_point(laszip_POINTER pointer, const laszip_point* point)
{
  static_cast<laszip_shim::Handle*>(pointer)->point = *point;
  return 0;
}
inline laszip_I32
laszip_set_coordinates(laszip_POINTER pointer, const laszip_F64* coordinates)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  h->point.X = laszip_shim::quantize((coordinates[0] - h->header.x_offset) / h->header.x_scale_factor);
  h->point.Y = laszip_shim::quantize((coordinates[1] - h->header.y_offset) / h->header.y_scale_factor);
  h->point.Z = laszip_shim::quantize((coordinates[2] - h->header.z_offset) / h->header.z_scale_factor);
  return 0;
}
inline laszip_I32
laszip_open_writer(laszip_POINTER pointer, const laszip_CHAR* file_name, laszip_BOOL compress)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  h->file = &laszip_shim::files()[file_name];
  h->file->header = h->header;
  h->file->points.clear();
  h->writing = true;
  return 0;
}
inline laszip_I32
laszip_write_point(laszip_POINTER pointer)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  if (!h->file)
    return 1;
  h->file->points.push_back(h->point);
  return 0;
}
inline laszip_I32
laszip_close_writer(laszip_POINTER pointer)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  h->file = nullptr;
  h->writing = false;
  return 0;
}
inline laszip_I32
laszip_open_reader(laszip_POINTER pointer, const laszip_CHAR* file_name, laszip_BOOL* is_compressed)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  auto it = laszip_shim::files().find(file_name);
  if (it == laszip_shim::files().end())
    return 1;
  h->file = &it->second;
  h->header = it->second.header;
  h->cursor = 0;
  if (is_compressed)
    *is_compressed = 0;
  return 0;
}
inline laszip_I32
laszip_seek_point(laszip_POINTER pointer, laszip_I64 index)
{
  static_cast<laszip_shim::Handle*>(pointer)->cursor = static_cast<size_t>(index);
  return 0;
}
inline laszip_I32
laszip_read_point(laszip_POINTER pointer)
{
  auto* h = static_cast<laszip_shim::Handle*>(pointer);
  if (!h->file || h->cursor >= h->file->points.size())
    return 1;
  h->point = h->file->points[h->cursor++];
  return 0;
}
inline laszip_I32
laszip_close_reader(laszip_POINTER pointer)
{
  static_cast<laszip_shim::Handle*>(pointer)->file = nullptr;
  return 0;
}
