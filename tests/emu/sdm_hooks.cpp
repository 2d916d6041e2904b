// test-only C hooks around the product's host-side .sdm / image code (iris_lama_b200/csrc/sdm_io.cpp), so that it is exercised
// without a GPU: planes in, file / image out, and back.
#include <cstring>
#include <string>

#include "../../iris_lama_b200/csrc/sdm_io.h"

using namespace lama_b200;

extern "C" {

int sdmtest_write_distance(const char* path, float resolution, uint32_t max_sqdist, uint32_t x0, uint32_t y0, int w, int h, const uint16_t* sqdist,
                           const uint8_t* valid, const uint8_t* known, const int16_t* ox, const int16_t* oy, const uint8_t* queued)
{
    SdmWindow win; win.x0 = x0; win.y0 = y0; win.w = w; win.h = h;
    SdmFile f;
    sdm_from_distance(win, resolution, max_sqdist, sqdist, valid, known, ox, oy, queued, f);
    std::string err;
    return sdm_write(path, f, err) ? 1 : 0;
}
int sdmtest_write_frequency(const char* path, float resolution, uint32_t x0, uint32_t y0, int w, int h, const uint16_t* occupied, const uint16_t* visited,
                            const uint8_t* known)
{
    SdmWindow win; win.x0 = x0; win.y0 = y0; win.w = w; win.h = h;
    SdmFile f;
    sdm_from_frequency(win, resolution, occupied, visited, known, f);
    std::string err;
    return sdm_write(path, f, err) ? 1 : 0;
}
// reads a distance-map file; win_out = {x0, y0, w, h}; planes may be NULL to query the window only
int sdmtest_read_distance(const char* path, uint32_t* win_out, uint32_t* max_sqdist, uint16_t* sqdist, uint8_t* valid, uint8_t* known, int16_t* ox, int16_t* oy,
                          uint8_t* queued)
{
    SdmFile f;
    std::string err;
    if (!sdm_read(path, sizeof(SdmDistanceCell), 4, f, err)) return 0;
    std::memcpy(max_sqdist, f.params.data(), 4);
    SdmWindow win;
    if (!sdm_window_of(f, win)) return 0;
    win_out[0] = win.x0; win_out[1] = win.y0; win_out[2] = (uint32_t)win.w; win_out[3] = (uint32_t)win.h;
    if (sqdist) sdm_to_distance(f, win, sqdist, valid, known, ox, oy, queued);
    return 1;
}
void sdmtest_distance_image(uint32_t x0, uint32_t y0, int w, int h, const uint16_t* sqdist, const uint8_t* valid, const uint8_t* known, uint32_t max_sqdist,
                            double resolution, uint8_t* out)
{
    SdmWindow win; win.x0 = x0; win.y0 = y0; win.w = w; win.h = h;
    sdm_distance_image(win, sqdist, valid, known, max_sqdist, resolution, out);
}
void sdmtest_frequency_image(uint32_t x0, uint32_t y0, int w, int h, const uint16_t* occupied, const uint16_t* visited, const uint8_t* known, uint8_t* out)
{
    SdmWindow win; win.x0 = x0; win.y0 = y0; win.w = w; win.h = h;
    sdm_occupancy_image_frequency(win, occupied, visited, known, out);
}
}
