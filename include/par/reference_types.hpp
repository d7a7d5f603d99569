// reference_types.hpp — C++ host-side mirror of the reference's scene interface, for hosts
// that want to keep writing against the reference's names:
//   Point<T>, AABB            /root/reference/src/alternative.cpp:12-38   (layout identical)
//   Entities<N>::insert/size  alternative.cpp:92-114
//   Light                     alternative.cpp:619-622
//   Color, Vector, Pixel      sprites.hpp:5-58
//   Sprite, make_tile_floor   sprites.hpp:67-71, 73-364
// plus par::FrameRenderer, which replaces the frame-loop body alternative.cpp:689-760 with
// calls into the C ABI (par/par.h).  Header only; no CUDA types.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "par/par.h"

namespace par {

template <typename T>
struct Point {
    T x, y, z;
};

struct alignas(16) AABB {  // alternative.cpp:35-38
    Point<short> position;
    Point<short> extent;
};
static_assert(sizeof(AABB) == sizeof(par_aabb), "AABB layout");

struct Color {  // sprites.hpp:5-6
    unsigned char red, green, blue, alpha;
};
static_assert(sizeof(Color) == sizeof(par_color), "Color layout");

template <typename T = float>
struct Vector {  // sprites.hpp:20-22
    T x, y, z;
};

struct Pixel {  // sprites.hpp:53-58
    Vector<float> normal;
    Color color;
    int y, z;
    int entity_index;
};
static_assert(sizeof(Pixel) == sizeof(par_pixel), "Pixel layout");

struct Sprite {  // sprites.hpp:67-71
    std::array<int, PAR_SPRITE_TEXELS> color;
    std::array<int, PAR_SPRITE_TEXELS> depth;
    std::array<Vector<float>, PAR_SPRITE_TEXELS> normal;
};
static_assert(sizeof(Sprite) == sizeof(par_sprite), "Sprite layout");

struct Light {  // alternative.cpp:619-622
    short x, y, z;
    short radius = 10;
};
static_assert(sizeof(Light) == sizeof(par_light), "Light layout");

inline Sprite make_tile_floor() {  // sprites.hpp:73-364
    Sprite s;
    par_sprite_tile_floor(reinterpret_cast<par_sprite*>(&s));
    return s;
}

// Entities, alternative.cpp:92-114.  The reference stores one 16 000-byte Sprite copy per
// entity (2.6 GB for the default scene); here sprites live once in a pool and every entity
// keeps an index, which is also what the device wants.  insert() reproduces the reference's
// behaviour of ignoring the caller's sprite and using tile_single (quirk Q1,
// alternative.cpp:105-108); insert_with_sprite() is the lifted version.
template <int entity_count = 0>
struct Entities {
    struct Entity {
        AABB aabb;
        Sprite sprite;
    };

    std::vector<AABB> aabbs;
    std::vector<int32_t> sprite_ids;
    std::vector<Sprite> sprite_pool{make_tile_floor()};  // entry 0 = tile_single
    int last_entity_index = 0;

    void insert(const Entity& entity) { push(entity.aabb, 0); }
    void insert(const AABB& aabb) { push(aabb, 0); }

    void insert_with_sprite(const Entity& entity) {
        int id = -1;
        for (size_t i = 0; i < sprite_pool.size() && id < 0; i++)
            if (std::memcmp(&sprite_pool[i], &entity.sprite, sizeof(Sprite)) == 0) id = static_cast<int>(i);
        if (id < 0) {
            sprite_pool.push_back(entity.sprite);
            id = static_cast<int>(sprite_pool.size()) - 1;
            atlas_dirty = true;
        }
        push(entity.aabb, id);
    }

    int size() const { return last_entity_index; }
    bool atlas_dirty = true;

  private:
    void push(const AABB& aabb, int sprite) {
        aabbs.push_back(aabb);
        sprite_ids.push_back(sprite);
        last_entity_index += 1;
    }
};

// The default scene of alternative.cpp:519-599 through Entities::insert.
template <int N>
inline void make_default_scene(Entities<N>& e) {
    const int n = par_scene_default(nullptr, 0);
    std::vector<par_aabb> boxes(static_cast<size_t>(n));
    par_scene_default(boxes.data(), n);
    for (const par_aabb& b : boxes) e.insert(AABB{{b.px, b.py, b.pz}, {b.ex, b.ey, b.ez}});
}

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// Replaces the frame-loop body alternative.cpp:689-760.
class FrameRenderer {
  public:
    FrameRenderer(int view_width, int view_height, int view_length, int device = 0) {
        par_config cfg{};
        cfg.width = view_width;
        cfg.height = view_height;
        cfg.length = view_length;
        cfg.device = device;
        check(par_create(&ctx_, &cfg));
    }
    ~FrameRenderer() {
        par_destroy(ctx_);
        for (Staging& st : staging_) {
            par_free_host(st.aabbs);
            par_free_host(st.sprite_ids);
        }
    }
    FrameRenderer(const FrameRenderer&) = delete;
    FrameRenderer& operator=(const FrameRenderer&) = delete;

    // One frame: memset + count_entities_in_bins + trace_hash_for_pixel + shading loop.
    // p_texture: view_width*view_height Color (alternative.cpp:515); p_pixel_buffer: optional
    // G-buffer (alternative.cpp:511).
    template <int N>
    void render_frame(Entities<N>& entities, const std::vector<Light>& lights, Color* p_texture,
                      Pixel* p_pixel_buffer = nullptr, par_stats* stats = nullptr) {
        update_atlas(entities);
        check(par_set_scene(ctx_, reinterpret_cast<const par_aabb*>(entities.aabbs.data()),
                            entities.sprite_ids.data(), entities.size()));
        check(par_render(ctx_, reinterpret_cast<const par_light*>(lights.data()),
                         static_cast<int>(lights.size()), reinterpret_cast<par_color*>(p_texture),
                         reinterpret_cast<par_pixel*>(p_pixel_buffer), stats));
        resident_entities_ = entities.size();
    }

    // Throughput form of render_frame: up to two frames in flight.  submit_frame snapshots the
    // entities into page-locked staging (so the caller may move them right away) and returns;
    // wait_frame blocks until the OLDEST submitted frame is complete in the p_texture it was given
    // (allocate it with alloc_frame(): page-locked memory keeps the readback asynchronous).
    template <int N>
    void submit_frame(Entities<N>& entities, const std::vector<Light>& lights, Color* p_texture) {
        if (in_flight_ >= 2)  // before touching the staging buffers: the older frame may still be uploading from them
            throw Error(PAR_ERR_STATE, "FrameRenderer::submit_frame: two frames in flight, call wait_frame first");
        update_atlas(entities);
        Staging& st = staging_[next_ & 1];
        const size_t n = static_cast<size_t>(entities.size());
        if (n > st.capacity) {
            par_free_host(st.aabbs);
            par_free_host(st.sprite_ids);
            st.capacity = n + n / 8 + 64;
            st.aabbs = static_cast<par_aabb*>(par_alloc_host(st.capacity * sizeof(par_aabb)));
            st.sprite_ids = static_cast<int32_t*>(par_alloc_host(st.capacity * sizeof(int32_t)));
            if (!st.aabbs || !st.sprite_ids) throw Error(PAR_ERR_OUT_OF_MEMORY, par_last_error());
        }
        std::memcpy(st.aabbs, entities.aabbs.data(), n * sizeof(par_aabb));
        std::memcpy(st.sprite_ids, entities.sprite_ids.data(), n * sizeof(int32_t));
        check(par_submit_frame(ctx_, st.aabbs, st.sprite_ids, static_cast<int>(n),
                               reinterpret_cast<const par_light*>(lights.data()), static_cast<int>(lights.size()),
                               reinterpret_cast<par_color*>(p_texture)));
        next_++;
        in_flight_++;
        resident_entities_ = entities.size();
    }
    // The same when only entities [first, first + count) moved since the previous frame — what the
    // reference's key handler does to entity 0 (alternative.cpp:641-660): nothing but those records
    // travels to the GPU and only the bins they touch are rebuilt (par_submit_update).  Falls back to
    // a full submit_frame when no scene of this size is resident yet.
    template <int N>
    void submit_frame_moved(Entities<N>& entities, int first, int count, const std::vector<Light>& lights,
                            Color* p_texture) {
        if (resident_entities_ != entities.size() || entities.atlas_dirty) {
            submit_frame(entities, lights, p_texture);
            return;
        }
        if (in_flight_ >= 2)
            throw Error(PAR_ERR_STATE, "FrameRenderer::submit_frame_moved: two frames in flight, call wait_frame first");
        check(par_submit_update(ctx_, first, count, reinterpret_cast<const par_aabb*>(entities.aabbs.data() + first),
                                entities.sprite_ids.data() + first, reinterpret_cast<const par_light*>(lights.data()),
                                static_cast<int>(lights.size()), reinterpret_cast<par_color*>(p_texture)));
        in_flight_++;
    }
    void wait_frame(par_stats* stats = nullptr) {
        check(par_wait_frame(ctx_, stats));
        in_flight_--;
    }
    // Row pitch of the host frames handed to render_frame / submit_frame (alternative.cpp:774-783: the
    // locked SDL texture has its own pitch); 0 = packed.
    void set_output_pitch(size_t pitch_bytes) { check(par_set_output_pitch(ctx_, pitch_bytes)); }

    // The record under the mouse cursor (mouse_pixel, alternative.cpp:380-382) of the frame
    // most recently completed, for the debug overlay.
    void set_cursor(int x, int y) { check(par_set_cursor(ctx_, x, y)); }
    Pixel cursor_pixel() {
        Pixel p;
        check(par_cursor_pixel(ctx_, reinterpret_cast<par_pixel*>(&p)));
        return p;
    }

    static Color* alloc_frame(int view_width, int view_height) {
        void* p = par_alloc_host(static_cast<size_t>(view_width) * view_height * sizeof(Color));
        if (!p) throw Error(PAR_ERR_OUT_OF_MEMORY, par_last_error());
        return static_cast<Color*>(p);
    }
    static void free_frame(Color* p) { par_free_host(p); }

    par_ctx* handle() { return ctx_; }

  private:
    struct Staging {
        par_aabb* aabbs = nullptr;
        int32_t* sprite_ids = nullptr;
        size_t capacity = 0;
    };
    template <int N>
    void update_atlas(Entities<N>& entities) {
        static const par_color palette[4] = {{100, 100, 100, 0}, {140, 140, 140, 0},
                                             {200, 200, 200, 0}, {240, 240, 240, 0}};  // sprites.hpp:60-65
        if (entities.atlas_dirty) {
            check(par_set_atlas(ctx_, reinterpret_cast<const par_sprite*>(entities.sprite_pool.data()),
                                static_cast<int>(entities.sprite_pool.size()), palette, 4));
            entities.atlas_dirty = false;
        }
    }
    static void check(int rc) {
        if (rc != PAR_OK) throw Error(rc, par_last_error());
    }
    par_ctx* ctx_ = nullptr;
    Staging staging_[2];
    unsigned next_ = 0;
    int in_flight_ = 0;
    int resident_entities_ = -1;  // size of the scene resident on the device (-1: none)
};

}  // namespace par
