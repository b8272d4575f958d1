// Scene front end: the reference's JSON constructor vocabulary (src/scene.rs:618-1408) -> an
// in-memory scene description -> EuclFlatScene tables (include/euclider_b200.h).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "euclider_b200.h"

namespace eucl {

struct ParsedScene {
    // owned storage behind the borrowed pointers of `flat`
    std::vector<EuclPrim> prims;
    std::vector<EuclNode> nodes;
    std::vector<EuclEntity> entities;
    std::vector<EuclMaterial> materials;
    std::vector<EuclTransform> transforms;
    std::vector<EuclExprOp> expr_ops;
    std::vector<EuclSurface> surfaces;
    std::vector<EuclColorOp> color_ops;
    std::vector<EuclMappedTexture> mapped_textures;
    std::vector<EuclTexture> textures;
    std::vector<uint8_t> texels;

    std::vector<std::string> texture_paths; // one per texture slot
    std::vector<std::vector<uint8_t>> texture_pixels;
    std::vector<bool> texture_set;

    EuclFlatScene flat{};
    void refresh_flat(); // re-point `flat` at the vectors, repack texels
};

// Mirrors Parser::parse::<Box<Environment>> (src/scene.rs:1466-1478).  On failure returns a
// negative EuclStatus (the ParserError variant) and a message.
int parse_scene(const std::string& json_text, std::unique_ptr<ParsedScene>* out, std::string* error);

// noise 0.4.1 PermutationTable::new(seed) (RECOLLECTION, see oracle/ASSUMPTIONS.md)
void perlin_permutation(uint32_t seed, uint8_t out[256]);

// palette 0.2.1 Hsv -> Rgb (RECOLLECTION), used by `Rgba::from_hsva` (src/scene.rs:662-667)
void hsv_to_rgb(double hue_degrees, double saturation, double value, double rgb[3]);

} // namespace eucl
