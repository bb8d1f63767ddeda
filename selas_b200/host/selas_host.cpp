// selas_host.cpp -- see selas_host.h.  Host-side glue only: everything on the hot path happens behind the C ABI.
#include "selas_host.h"

#include <sys/stat.h>

#include <cctype>
#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace Selas {

Error MakeError(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return Error(buf);
}

// ---------------------------------------------------------------------------------------------------------------
// A minimal JSON reader: objects of numbers / strings / bools / flat number arrays is all the material files use
// (Content/Materials/*.json; the reference reads them with rapidjson through UtilityLib/JsonUtilities).
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct JsonValue {
    enum Kind { kNull, kNumber, kString, kBool, kArray, kObject } kind = kNull;
    double number = 0;
    bool boolean = false;
    std::string string;
    std::vector<JsonValue> array;
    std::map<std::string, JsonValue> object;
};
struct JsonParser {
    const char* p;
    std::string error;
    void ws() {
        while (*p && isspace((unsigned char)*p)) ++p;
    }
    bool fail(const char* m) {
        if (error.empty()) error = m;
        return false;
    }
    bool parseString(std::string& out) {
        if (*p != '"') return fail("expected string");
        ++p;
        out.clear();
        while (*p && *p != '"') {
            if (*p == '\\' && p[1]) {
                ++p;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    default: out += *p;
                }
                ++p;
            } else out += *p++;
        }
        if (*p != '"') return fail("unterminated string");
        ++p;
        return true;
    }
    bool parseValue(JsonValue& v) {
        ws();
        if (*p == '{') {
            v.kind = JsonValue::kObject;
            ++p;
            ws();
            if (*p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string key;
                if (!parseString(key)) return false;
                ws();
                if (*p != ':') return fail("expected ':'");
                ++p;
                if (!parseValue(v.object[key])) return false;
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JsonValue::kArray;
            ++p;
            ws();
            if (*p == ']') { ++p; return true; }
            for (;;) {
                v.array.emplace_back();
                if (!parseValue(v.array.back())) return false;
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') {
            v.kind = JsonValue::kString;
            return parseString(v.string);
        }
        if (!strncmp(p, "true", 4)) { v.kind = JsonValue::kBool; v.boolean = true; p += 4; return true; }
        if (!strncmp(p, "false", 5)) { v.kind = JsonValue::kBool; v.boolean = false; p += 5; return true; }
        if (!strncmp(p, "null", 4)) { v.kind = JsonValue::kNull; p += 4; return true; }
        char* end = nullptr;
        double d = strtod(p, &end);
        if (end == p) return fail("unexpected character");
        v.kind = JsonValue::kNumber;
        v.number = d;
        p = end;
        return true;
    }
};

const char* kAttributes[SELAS_SCALAR_COUNT] = {"metallic", "specularTint", "roughness", "anisotropic", "sheen", "sheenTint", "clearcoat",
                                               "clearcoatGloss", "specTrans", "diffTrans", "flatness", "ior", "scatterDistance", "displacement"};
// ImportMaterial.cpp:36-52
const float kAttributeDefaults[SELAS_SCALAR_COUNT] = {0.3f, 0.0f, 0.5f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.5f, 0.0f, 0.0f};

bool equalsIgnoreCase(const std::string& a, const char* b) {
    size_t n = strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
    return true;
}
void readFloat3(const JsonValue& doc, const char* key, float* out, float d0, float d1, float d2) {
    out[0] = d0; out[1] = d1; out[2] = d2;
    auto it = doc.object.find(key);
    if (it == doc.object.end() || it->second.kind != JsonValue::kArray || it->second.array.size() < 3) return;
    for (int i = 0; i < 3; ++i)
        if (it->second.array[i].kind == JsonValue::kNumber) out[i] = (float)it->second.array[i].number;
}
}  // namespace

Error ImportMaterialFromString(const char* json, ImportedMaterialData* material) {
    JsonParser parser{json, ""};
    JsonValue doc;
    parser.ws();
    if (*parser.p == 0) doc.kind = JsonValue::kObject;  // an empty file behaves like "{}" (Default.json is "{ }")
    else if (!parser.parseValue(doc) || doc.kind != JsonValue::kObject)
        return Error_("material json: %s", parser.error.empty() ? "top-level value is not an object" : parser.error.c_str());

    auto str = [&](const char* key, std::string& out) {
        auto it = doc.object.find(key);
        out = (it != doc.object.end() && it->second.kind == JsonValue::kString) ? it->second.string : std::string();
    };
    auto isString = [&](const char* key) {
        auto it = doc.object.find(key);
        return it != doc.object.end() && it->second.kind == JsonValue::kString;
    };
    auto boolean = [&](const char* key, bool dflt) {
        auto it = doc.object.find(key);
        return (it != doc.object.end() && it->second.kind == JsonValue::kBool) ? it->second.boolean : dflt;
    };
    str("shaderName", material->shaderName);
    material->baseColorTexture.clear();
    if (isString("baseColor")) {  // ImportMaterial.cpp:63-67
        str("baseColor", material->baseColorTexture);
        material->baseColor[0] = material->baseColor[1] = material->baseColor[2] = 0.6f;
    } else {
        readFloat3(doc, "baseColor", material->baseColor, 0.6f, 0.6f, 0.6f);
    }
    readFloat3(doc, "transmittanceColor", material->transmittanceColor, 0.f, 0.f, 0.f);
    for (int i = 0; i < SELAS_SCALAR_COUNT; ++i) {
        material->scalarAttributeTextures[i].clear();
        material->scalarAttributes[i] = kAttributeDefaults[i];
        if (isString(kAttributes[i])) str(kAttributes[i], material->scalarAttributeTextures[i]);
        else {
            auto it = doc.object.find(kAttributes[i]);
            if (it != doc.object.end() && it->second.kind == JsonValue::kNumber) material->scalarAttributes[i] = (float)it->second.number;
        }
    }
    material->alphaTested = boolean("alphaTesting", false);
    material->invertDisplacement = boolean("invertDisplacement", false);
    material->usesPtex = boolean("usesPtex", false);
    return Success_;
}

Error ImportMaterial(const char* filepath, ImportedMaterialData* material) {
    FILE* fp = fopen(filepath, "rb");
    if (!fp) return Error_("Failed to open material file: %s", filepath);
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
    fclose(fp);
    return ImportMaterialFromString(text.c_str(), material);
}

void BuildMaterial(const ImportedMaterialData& imported, SelasMaterial& material) {  // BuildModel.cpp:22-66
    memset(&material, 0, sizeof(material));
    if (imported.alphaTested) material.flags |= SELAS_MAT_ALPHA_TESTED;
    if (imported.invertDisplacement) material.flags |= SELAS_MAT_INVERT_DISPLACEMENT;
    if (imported.usesPtex) material.flags |= SELAS_MAT_USES_PTEX;
    static const char* shaderNames[] = {"DisneyThin", "DisneySolid", "DeltaTransparent"};
    material.shader = SELAS_SHADER_DISNEY_SOLID;
    for (uint32_t i = 0; i < 3; ++i)
        if (equalsIgnoreCase(imported.shaderName, shaderNames[i])) material.shader = i;
    memcpy(material.baseColor, imported.baseColor, sizeof(material.baseColor));
    memcpy(material.transmittanceColor, imported.transmittanceColor, sizeof(material.transmittanceColor));
    for (int i = 0; i < SELAS_SCALAR_COUNT; ++i) material.scalars[i] = imported.scalarAttributes[i];
    if (material.shader == SELAS_SHADER_DISNEY_SOLID && (material.scalars[SELAS_DIFFUSE_TRANS] > 0.0f || material.scalars[SELAS_SPEC_TRANS] > 0.0f))
        material.flags |= SELAS_MAT_TRANSPARENT;
}

// ---------------------------------------------------------------------------------------------------------------
// film
// ---------------------------------------------------------------------------------------------------------------
void FrameBuffer_Initialize(Framebuffer* frame, uint32_t width, uint32_t height, uint32_t layerCount) {
    frame->width = width;
    frame->height = height;
    frame->layerCount = layerCount;
    frame->buffers.assign(layerCount, std::vector<float3>((size_t)width * height, float3{0.f, 0.f, 0.f}));
}
void FrameBuffer_Shutdown(Framebuffer* frame) {
    frame->buffers.clear();
    frame->width = frame->height = frame->layerCount = 0;
}
void FrameBuffer_Scale(Framebuffer* frame, float scale) {  // Framebuffer.cpp:62-70
    for (auto& layer : frame->buffers)
        for (float3& p : layer) {
            p.x = p.x * scale;
            p.y = p.y * scale;
            p.z = p.z * scale;
        }
}

// Radiance RGBE as FrameBuffer_Save's writer produces it (StbImageWrite(..., HDR, ...) = stbi_write_hdr, TextureLib/StbImage.cpp:76-78):
// the same header lines, pixels as (mantissa r g b, shared exponent + 128) truncated -- not rounded -- from frexp of the largest channel,
// black below 1e-32; scanlines of 8 <= width < 32768 in the format's adaptive run-length form (2 2 hi lo, then the four components one
// after the other: literal chunks of <= 128 bytes, runs of >= 3 equal bytes as (128 + n <= 255, value)), others flat.  tests/test_host.py
// compares the files byte for byte with the reference writer's.
static void rgbeOf(const float3& p, unsigned char* o) {
    const float r = p.x, g = p.y, b = p.z;
    const float m = r > g ? (r > b ? r : b) : (g > b ? g : b);
    if (!(m >= 1e-32f)) {
        o[0] = o[1] = o[2] = o[3] = 0;
        return;
    }
    int e;
    const float norm = frexpf(m, &e) * 256.0f / m;
    o[0] = (unsigned char)(r * norm);
    o[1] = (unsigned char)(g * norm);
    o[2] = (unsigned char)(b * norm);
    o[3] = (unsigned char)(e + 128);
}

static void rleComponent(const unsigned char* c, uint32_t w, std::vector<unsigned char>& out) {
    uint32_t x = 0;
    while (x < w) {
        uint32_t r = x;  // start of the next run of three equal bytes, or w
        while (r + 2 < w && !(c[r] == c[r + 1] && c[r] == c[r + 2])) ++r;
        if (r + 2 >= w) r = w;
        while (x < r) {  // literals up to the run
            const uint32_t n = std::min<uint32_t>(128, r - x);
            out.push_back((unsigned char)n);
            out.insert(out.end(), c + x, c + x + n);
            x += n;
        }
        if (r + 2 < w) {  // the run itself
            while (r < w && c[r] == c[x]) ++r;
            while (x < r) {
                const uint32_t n = std::min<uint32_t>(127, r - x);
                out.push_back((unsigned char)(128 + n));
                out.push_back(c[x]);
                x += n;
            }
        }
    }
}

Error WriteRadianceHdr(const char* path, uint32_t width, uint32_t height, const float3* pixels) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return Error_("Failed to open %s for writing", path);
    // the header the reference's files carry (stb's comment line included, so the two writers' files compare equal)
    fprintf(fp, "#?RADIANCE\n# Written by stb_image_write.h\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y %u +X %u\n", height, width);
    std::vector<unsigned char> row((size_t)width * 4), planes((size_t)width * 4), out;
    const bool rle = width >= 8 && width < 32768;
    for (uint32_t y = 0; y < height; ++y) {
        for (uint32_t x = 0; x < width; ++x) rgbeOf(pixels[(size_t)y * width + x], &row[(size_t)x * 4]);
        out.clear();
        if (rle) {
            out.push_back(2);
            out.push_back(2);
            out.push_back((unsigned char)(width >> 8));
            out.push_back((unsigned char)(width & 0xff));
            for (int c = 0; c < 4; ++c) {
                for (uint32_t x = 0; x < width; ++x) planes[(size_t)c * width + x] = row[(size_t)x * 4 + c];
                rleComponent(&planes[(size_t)c * width], width, out);
            }
        } else {
            out = row;
        }
        if (fwrite(out.data(), 1, out.size(), fp) != out.size()) {
            fclose(fp);
            return Error_("short write on %s", path);
        }
    }
    fclose(fp);
    return Success_;
}

Error FrameBuffer_Save(Framebuffer* frame, const char* name) {  // Framebuffer.cpp:43-59: _Images/<name>_<layer>.hdr
    mkdir("_Images", 0755);
    for (uint32_t l = 0; l < frame->layerCount; ++l) {
        char path[1024];
        snprintf(path, sizeof(path), "_Images/%s_%u.hdr", name, l);
        ReturnError_(WriteRadianceHdr(path, frame->width, frame->height, frame->buffers[l].data()));
    }
    return Success_;
}

// ---------------------------------------------------------------------------------------------------------------
// camera / scene
// ---------------------------------------------------------------------------------------------------------------
void InitializeRayCastCamera(const CameraSettings& settings, uint32_t width, uint32_t height, RayCastCameraSettings& camera) {
    SelasCameraSettings st;
    memcpy(st.position, settings.position, sizeof(st.position));
    st.fovDegrees = settings.fovDegrees;
    memcpy(st.lookAt, settings.lookAt, sizeof(st.lookAt));
    st.znear = settings.znear;
    memcpy(st.up, settings.up, sizeof(st.up));
    st.zfar = settings.zfar;
    selas_cuda_camera_init(&st, width, height, &camera);
}

// SceneResource::aaBox / boundingSphere from the meshes and the instance list (no device work).
void CalculateSceneBoundingBox(SceneResource* scene) {
    // SceneResource::aaBox / boundingSphere (SceneResource.h:77-78), CalculateSceneBoundingBox (SceneResource.cpp:194-203): every
    // instance contributes the eight corners of its mesh's object-space box taken through localToWorld (IncludeBox,
    // AxisAlignedBox.cpp:44-65: tmin + combinations of the transformed edge vectors).  Scenes without an instance list are one
    // identity instance over all meshes: the raw vertices.
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    std::vector<float> meshBox((scene->meshes.size() + scene->curves.size()) * 6);
    for (size_t ci = 0; ci < scene->curves.size(); ++ci) {  // curve geometries follow the meshes: control vertices grown by their radii
        const SceneCurve& c = scene->curves[ci];
        float* b = &meshBox[(scene->meshes.size() + ci) * 6];
        for (int k = 0; k < 3; ++k) { b[k] = 3.402823466e+38f; b[3 + k] = -3.402823466e+38f; }
        for (size_t v = 0; v + 3 < c.vertices.size(); v += 4)
            for (int k = 0; k < 3; ++k) {
                b[k] = std::min(b[k], c.vertices[v + k] - std::fabs(c.vertices[v + 3]));
                b[3 + k] = std::max(b[3 + k], c.vertices[v + k] + std::fabs(c.vertices[v + 3]));
            }
    }
    for (size_t mi = 0; mi < scene->meshes.size(); ++mi) {
        const SceneMesh& m = scene->meshes[mi];
        float* b = &meshBox[mi * 6];
        for (int k = 0; k < 3; ++k) { b[k] = 3.402823466e+38f; b[3 + k] = -3.402823466e+38f; }
        for (size_t v = 0; v + 2 < m.positions.size(); v += 3)
            for (int k = 0; k < 3; ++k) {
                b[k] = std::min(b[k], m.positions[v + k]);
                b[3 + k] = std::max(b[3 + k], m.positions[v + k]);
            }
    }
    auto includePoint = [&](const float* p) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
    };
    if (scene->instances.empty()) {
        for (size_t mi = 0; mi < scene->meshes.size() + scene->curves.size(); ++mi)
            if (meshBox[mi * 6] <= meshBox[mi * 6 + 3]) {
                includePoint(&meshBox[mi * 6]);
                includePoint(&meshBox[mi * 6 + 3]);
            }
    } else {
        for (const SelasInstanceDesc& inst : scene->instances) {
            if (inst.meshIndex >= scene->meshes.size() + scene->curves.size()) continue;
            const float* b = &meshBox[(size_t)inst.meshIndex * 6];
            if (!(b[0] <= b[3])) continue;  // empty mesh
            const float* M = inst.localToWorld;  // row-major, row vectors, translation in row 3
            float tmin[3], e[3][3];
            for (int c = 0; c < 3; ++c) {
                tmin[c] = b[0] * M[0 + c] + b[1] * M[4 + c] + b[2] * M[8 + c] + M[12 + c];
                for (int a = 0; a < 3; ++a) e[a][c] = (b[3 + a] - b[a]) * M[4 * a + c];
            }
            for (int corner = 0; corner < 8; ++corner) {
                float p[3];
                for (int c = 0; c < 3; ++c)
                    p[c] = tmin[c] + ((corner & 1) ? e[0][c] : 0.0f) + ((corner & 2) ? e[1][c] : 0.0f) + ((corner & 4) ? e[2][c] : 0.0f);
                includePoint(p);
            }
        }
    }
    if (!(lo[0] <= hi[0]))
        for (int k = 0; k < 3; ++k) lo[k] = hi[k] = 0.0f;
    double diag2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        scene->aaBox[k] = lo[k];
        scene->aaBox[3 + k] = hi[k];
        scene->boundingSphere[k] = (float)(0.5 * ((double)lo[k] + (double)hi[k]));
        diag2 += ((double)hi[k] - (double)lo[k]) * ((double)hi[k] - (double)lo[k]);
    }
    scene->boundingSphere[3] = (float)(0.5 * sqrt(diag2));
}

Error InitializeSceneResource(SceneResource* scene, int device) {
    if (selas_cuda_context_create(device, &scene->gpu) != SELAS_OK) return Error_("%s", selas_cuda_last_error(nullptr));
    // TextureCache::LoadTexturePtex (TextureCache.cpp:99-126) for the meshes that name a Ptex file: one texture per distinct path
    std::vector<std::pair<std::string, SelasPtex*>> loadedPtex;
    std::vector<SelasPtex*> meshPtex(scene->meshes.size(), nullptr);
    for (size_t i = 0; i < scene->meshes.size(); ++i) {
        const SceneMesh& m = scene->meshes[i];
        meshPtex[i] = m.ptex;
        if (m.ptex || m.ptexPath.empty()) continue;
        for (const auto& lp : loadedPtex)
            if (lp.first == m.ptexPath) meshPtex[i] = lp.second;
        if (meshPtex[i]) continue;
        SelasPtex* tex = nullptr;
        if (selas_cuda_ptex_load(scene->gpu, m.ptexPath.c_str(), &tex) != SELAS_OK) {
            Error e = Error_("%s", selas_cuda_last_error(scene->gpu));
            ShutdownSceneResource(scene);
            return e;
        }
        scene->ownedPtex.push_back(tex);
        loadedPtex.emplace_back(m.ptexPath, tex);
        meshPtex[i] = tex;
    }
    std::vector<SelasMeshDesc> meshes(scene->meshes.size());
    for (size_t i = 0; i < meshes.size(); ++i) {
        const SceneMesh& m = scene->meshes[i];
        SelasMeshDesc& d = meshes[i];
        memset(&d, 0, sizeof(d));
        d.positions = m.positions.data();
        d.vertexCount = (uint32_t)(m.positions.size() / 3);
        d.indices = m.indices.data();
        d.indicesPerFace = m.indicesPerFace;
        d.faceCount = (uint32_t)(m.indices.size() / m.indicesPerFace);
        d.normals = m.normals.empty() ? nullptr : m.normals.data();
        d.tangents = m.tangents.empty() ? nullptr : m.tangents.data();
        d.uvs = m.uvs.empty() ? nullptr : m.uvs.data();
        d.materialIndex = m.materialIndex;
        d.lightSetIndex = m.lightSetIndex;
        d.ptex = meshPtex[i];
        d.baseColorTexture = m.baseColorTexture;
    }
    CalculateSceneBoundingBox(scene);
    std::vector<SelasCurveDesc> curves(scene->curves.size());
    for (size_t i = 0; i < curves.size(); ++i) {
        const SceneCurve& c = scene->curves[i];
        curves[i].vertices = c.vertices.data();
        curves[i].vertexCount = (uint32_t)(c.vertices.size() / 4);
        curves[i].indices = c.indices.data();
        curves[i].segmentCount = (uint32_t)c.indices.size();
        curves[i].materialIndex = c.materialIndex;
        curves[i].lightSetIndex = c.lightSetIndex;
    }
    SelasSceneDesc desc;
    memset(&desc, 0, sizeof(desc));
    desc.meshes = meshes.data();
    desc.meshCount = (uint32_t)meshes.size();
    desc.curves = curves.empty() ? nullptr : curves.data();
    desc.curveCount = (uint32_t)curves.size();
    desc.instances = scene->instances.empty() ? nullptr : scene->instances.data();
    desc.instanceCount = (uint32_t)scene->instances.size();
    desc.materials = scene->materials.data();
    desc.materialCount = (uint32_t)scene->materials.size();
    desc.lights = scene->lights.data();
    desc.lightCount = (uint32_t)scene->lights.size();
    desc.lightSets = scene->lightSets.data();
    desc.lightSetCount = (uint32_t)scene->lightSets.size();
    memcpy(desc.backgroundIntensity, scene->backgroundIntensity, sizeof(desc.backgroundIntensity));
    desc.ibl = scene->ibl;
    desc.flags = scene->flags;
    if (selas_cuda_scene_create(scene->gpu, &desc, &scene->gpuScene) != SELAS_OK) {
        Error e = Error_("%s", selas_cuda_last_error(scene->gpu));
        ShutdownSceneResource(scene);
        return e;
    }
    return Success_;
}

void ShutdownSceneResource(SceneResource* scene) {
    for (TextureResource* t : scene->textures) delete t;
    scene->textures.clear();
    scene->textureNames.clear();
    if (scene->gpuScene) selas_cuda_scene_destroy(scene->gpuScene);
    for (SelasPtex* t : scene->ownedPtex) selas_cuda_ptex_destroy(t);
    scene->ownedPtex.clear();
    if (scene->gpu) selas_cuda_context_destroy(scene->gpu);
    scene->gpuScene = nullptr;
    scene->gpu = nullptr;
}

namespace {
void addQuad(SceneMesh& m, const float a[3], const float b[3], const float c[3], const float d[3]) {
    uint32_t base = (uint32_t)(m.positions.size() / 3);
    const float* v[4] = {a, b, c, d};
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) m.positions.push_back(v[i][k]);
    for (uint32_t i = 0; i < 4; ++i) m.indices.push_back(base + i);
}
void addBlock(SceneMesh& m, float sx, float sy, float sz, float rotDeg, float cx, float cz) {
    const float c = cosf(rotDeg * 3.14159265358979f / 180.f), s = sinf(rotDeg * 3.14159265358979f / 180.f);
    auto P = [&](float x, float y, float z, float* o) {
        x -= sx / 2;
        z -= sz / 2;
        o[0] = cx + c * x + s * z;
        o[1] = y;
        o[2] = cz - s * x + c * z;
    };
    float v[8][3];
    P(0, 0, 0, v[0]); P(sx, 0, 0, v[1]); P(sx, 0, sz, v[2]); P(0, 0, sz, v[3]);
    P(0, sy, 0, v[4]); P(sx, sy, 0, v[5]); P(sx, sy, sz, v[6]); P(0, sy, sz, v[7]);
    addQuad(m, v[4], v[7], v[6], v[5]);
    addQuad(m, v[0], v[4], v[5], v[1]);
    addQuad(m, v[1], v[5], v[6], v[2]);
    addQuad(m, v[2], v[6], v[7], v[3]);
    addQuad(m, v[3], v[7], v[4], v[0]);
}
SelasMaterial makeMaterial(const char* json) {
    ImportedMaterialData imp;
    ImportMaterialFromString(json, &imp);
    SelasMaterial m;
    BuildMaterial(imp, m);
    return m;
}
}  // namespace

// BASELINE config 1; identical to selas_b200/scenes.py:cornell_box.
void MakeCornellBoxScene(SceneResource* scene, float copperRoughness) {
    const float X = 556.0f, Y = 548.8f, Z = 559.2f;
    scene->meshes.assign(4, SceneMesh());
    for (auto& m : scene->meshes) m.indicesPerFace = 4;
    SceneMesh& white = scene->meshes[0];
    {
        float a[3] = {0, 0, 0}, b[3] = {0, 0, Z}, c[3] = {X, 0, Z}, d[3] = {X, 0, 0};
        addQuad(white, a, b, c, d);
        float e[3] = {0, Y, 0}, f[3] = {X, Y, 0}, g[3] = {X, Y, Z}, h[3] = {0, Y, Z};
        addQuad(white, e, f, g, h);
        float i[3] = {0, 0, Z}, j[3] = {0, Y, Z}, k[3] = {X, Y, Z}, l[3] = {X, 0, Z};
        addQuad(white, i, j, k, l);
        addBlock(white, 165, 330, 165, 17.0f, 368.0f, 351.0f);
    }
    {
        float a[3] = {X, 0, 0}, b[3] = {X, 0, Z}, c[3] = {X, Y, Z}, d[3] = {X, Y, 0};
        addQuad(scene->meshes[1], a, b, c, d);
        float e[3] = {0, 0, 0}, f[3] = {0, Y, 0}, g[3] = {0, Y, Z}, h[3] = {0, 0, Z};
        addQuad(scene->meshes[2], e, f, g, h);
    }
    addBlock(scene->meshes[3], 165, 165, 165, -18.0f, 185.0f, 169.0f);
    for (uint32_t i = 0; i < 4; ++i) scene->meshes[i].materialIndex = i;
    const char* rough = "{\"shaderName\":\"DisneySolid\",\"baseColor\":[%g,%g,%g],\"ior\":1.2,\"metallic\":0.0,\"specularTint\":0.0,"
                        "\"roughness\":0.8,\"flatness\":0.5}";
    char buf[512];
    const float colors[3][3] = {{0.725f, 0.71f, 0.68f}, {0.63f, 0.065f, 0.05f}, {0.14f, 0.45f, 0.091f}};
    scene->materials.clear();
    for (int i = 0; i < 3; ++i) {
        snprintf(buf, sizeof(buf), rough, colors[i][0], colors[i][1], colors[i][2]);
        scene->materials.push_back(makeMaterial(buf));
    }
    snprintf(buf, sizeof(buf),
             "{\"shaderName\":\"DisneySolid\",\"baseColor\":[0.955008,0.637427,0.538163],\"ior\":1.5,\"metallic\":1.0,\"specularTint\":0.3,"
             "\"roughness\":%g,\"clearcoatGloss\":1.0}", copperRoughness);
    scene->materials.push_back(makeMaterial(buf));
    SelasLight L;
    memset(&L, 0, sizeof(L));
    const float pos[3] = {278.0f, 548.7f, 279.5f}, dir[3] = {0, -1, 0}, lx[3] = {130, 0, 0}, lz[3] = {0, 0, 105}, rad[3] = {17, 12, 4};
    memcpy(L.position, pos, 12); memcpy(L.direction, dir, 12); memcpy(L.x, lx, 12); memcpy(L.z, lz, 12); memcpy(L.radiance, rad, 12);
    scene->lights.assign(1, L);
    scene->lightSets.assign(1, SelasLightSetRange{0, 1});
    memset(scene->backgroundIntensity, 0, sizeof(scene->backgroundIntensity));
}

void SetupCornellCamera(uint32_t width, uint32_t height, RayCastCameraSettings& camera) {
    CameraSettings cs = {{278.0f, 273.0f, -800.0f}, 37.0f, {278.0f, 273.0f, 0.0f}, 0.1f, {0.0f, 1.0f, 0.0f}, 10000.0f};
    InitializeRayCastCamera(cs, width, height, camera);
}

// ---------------------------------------------------------------------------------------------------------------
// integrator entry point
// ---------------------------------------------------------------------------------------------------------------
namespace PathTracer {
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName, const PathTracerSettings& settings,
                    SelasRenderStats* stats, Framebuffer* keep) {
    if (!scene->gpuScene) return Error_("PathTracer::GenerateImage: scene is not initialised on a GPU (there is no CPU fallback)");
    const uint32_t width = (uint32_t)camera.width, height = (uint32_t)camera.height;
    Framebuffer frame;
    FrameBuffer_Initialize(&frame, width, height, settings.layerCount);
    SelasRenderSettings rs;
    memset(&rs, 0, sizeof(rs));
    rs.width = width;
    rs.height = height;
    rs.samplesPerPixel = settings.pathsPerPixel;
    rs.maxPathLength = settings.maxBounceCount;
    rs.rrStartBounce = settings.rrStartBounce;
    rs.layerCount = settings.layerCount;
    rs.seed = settings.seed;
    std::vector<float*> layers(settings.layerCount);
    for (uint32_t l = 0; l < settings.layerCount; ++l) layers[l] = &frame.buffers[l][0].x;
    if (selas_cuda_render_pathtracer(scene->gpuScene, &camera, &rs, layers.data(), stats) != SELAS_OK)
        return Error_("%s", selas_cuda_last_error(scene->gpu));
    FrameBuffer_Scale(&frame, 1.0f / (float)settings.pathsPerPixel);  // PathTracer.cpp:346
    ReturnError_(FrameBuffer_Save(&frame, imageName));
    if (keep) *keep = frame;
    FrameBuffer_Shutdown(&frame);
    return Success_;
}
}  // namespace PathTracer

namespace DeferredPathTracer {
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName, const DeferredPathTracerSettings& settings,
                    SelasRenderStats* stats, Framebuffer* keep) {
    if (!scene->gpuScene) return Error_("DeferredPathTracer::GenerateImage: scene is not initialised on a GPU (there is no CPU fallback)");
    const uint32_t width = (uint32_t)camera.width, height = (uint32_t)camera.height;
    const uint32_t spp = settings.samplesPerPixelX * settings.samplesPerPixelY;
    if (spp == 0 || settings.samplesPerPixelX > 0xffffu || settings.samplesPerPixelY > 0xffffu)
        return Error_("DeferredPathTracer::GenerateImage: SamplesPerPixelX_/Y_ out of range");
    Framebuffer frame;
    FrameBuffer_Initialize(&frame, width, height, 1);  // OutputLayers_, :42
    SelasRenderSettings rs;
    memset(&rs, 0, sizeof(rs));
    rs.width = width;
    rs.height = height;
    rs.samplesPerPixel = spp;
    rs.maxPathLength = settings.maxBounceCount;
    rs.layerCount = 1;
    rs.seed = settings.seed;
    rs.pixelSampler = settings.samplesPerPixelX | (settings.samplesPerPixelY << 16);  // JitteredCameraRay(..., scan, X, Y, index), :365-366
    rs.integrator = SELAS_INTEGRATOR_DEFERRED;
    float* layers[1] = {&frame.buffers[0][0].x};
    if (selas_cuda_render_pathtracer(scene->gpuScene, &camera, &rs, layers, stats) != SELAS_OK)
        return Error_("%s", selas_cuda_last_error(scene->gpu));
    FrameBuffer_Scale(&frame, 1.0f / (float)spp);  // :467
    ReturnError_(FrameBuffer_Save(&frame, imageName));
    if (keep) *keep = frame;
    FrameBuffer_Shutdown(&frame);
    return Success_;
}
}  // namespace DeferredPathTracer

// ---- image based light + config 5 scene -------------------------------------------------------------------------------
void ImageBasedLight::BuildDensityFunctions() {
    const float pi = 3.141592653589793f;
    marginal.assign(height, 0.f);
    conditional.assign((size_t)width * height, 0.f);
    std::vector<float> rowSum(height, 0.f);
    std::vector<float> inten(width);
    for (uint32_t y = 0; y < height; ++y) {
        float theta = ((float)y + 0.5f) * pi / (float)height;
        float st = sinf(theta);
        float run = 0.f;
        for (uint32_t x = 0; x < width; ++x) {
            const float* t = &lightData[3 * ((size_t)y * width + x)];
            float luma = t[0] * 0.299f + t[1] * 0.587f + t[2] * 0.114f;
            inten[x] = st * luma;
            run += inten[x];
        }
        rowSum[y] = run;
        if (run > 0.f) {
            float inv = 1.0f / run, c = 0.f;
            for (uint32_t x = 0; x < width; ++x) {
                c += inv * inten[x];
                conditional[(size_t)y * width + x] = c;
            }
        }
        conditional[(size_t)y * width + width - 1] = 1.0f;
    }
    float total = 0.f;
    for (uint32_t y = 0; y < height; ++y) total += rowSum[y];
    float c = 0.f;
    for (uint32_t y = 0; y < height; ++y) {
        c += total > 0.f ? rowSum[y] * (1.0f / total) : 0.f;
        marginal[y] = c;
    }
    marginal[height - 1] = 1.0f;
    memset(&desc, 0, sizeof(desc));
    desc.width = width;
    desc.height = height;
    desc.marginalDensityFunction = marginal.data();
    desc.conditionalDensityFunctions = conditional.data();
    desc.lightData = lightData.data();
    desc.missWidth = missWidth;
    desc.missHeight = missHeight;
    desc.missData = missData.data();
    desc.rotationRadians = rotationRadians;
    desc.exposureScale = exposureScale;
}

// ---- baked assets ---------------------------------------------------------------------------------------------------------
std::string AssetFilePath(const char* root, const char* type, uint64_t version, const char* name) {
    char buf[1024];
    snprintf(buf, sizeof(buf), "%s_Assets/%s/%llu/%s.bin", root, type, (unsigned long long)version, name);
    return std::string(buf);
}

namespace {
Error readWholeFile(const char* filepath, std::vector<uint8_t>& data) {  // File::ReadWholeFile, IoLib/File.cpp
    FILE* f = fopen(filepath, "rb");
    if (!f) return Error_("Failed to open file: %s", filepath);
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(size > 0 ? (size_t)size : 0);
    size_t got = data.empty() ? 0 : fread(data.data(), 1, data.size(), f);
    fclose(f);
    if (got != data.size()) return Error_("Failed to read file: %s", filepath);
    return Success_;
}
// a pointer field of an attached blob: the 8 bytes at `fieldOffset` hold the file offset of `size` bytes of data
Error blobPointer(const std::vector<uint8_t>& blob, size_t fieldOffset, uint64_t size, const uint8_t** out, const char* what) {
    uint64_t off = 0;
    memcpy(&off, blob.data() + fieldOffset, 8);
    if (off > blob.size() || size > blob.size() - off) return Error_("baked blob: %s lies outside the file (offset %llu, %llu bytes, file %zu)", what,
                                                                      (unsigned long long)off, (unsigned long long)size, blob.size());
    *out = blob.data() + off;
    return Success_;
}
}  // namespace

// ImageBasedLightResourceData as serialized (the struct itself, pointers = offsets; `uint` is 64 bits wide):
//   0 width u64 | 8 height u64 | 16 marginal* | 24 conditional* | 32 missWidth u64 | 40 missHeight u64 | 48 rotation f32 | 52 exposure f32
//   56 lightData* | 64 missData*                                                  (ImageBasedLightResource.h:16-35, .cpp:23-44)
Error ReadImageBasedLightResource(const char* filepath, ImageBasedLight* ibl) {
    std::vector<uint8_t> blob;
    ReturnError_(readWholeFile(filepath, blob));
    if (blob.size() < 72) return Error_("baked IBL blob %s: %zu bytes is shorter than the 72-byte resource header", filepath, blob.size());
    uint64_t w, h, mw, mh;
    memcpy(&w, blob.data() + 0, 8);
    memcpy(&h, blob.data() + 8, 8);
    memcpy(&mw, blob.data() + 32, 8);
    memcpy(&mh, blob.data() + 40, 8);
    if (w == 0 || h == 0 || w > 65536 || h > 65536 || mw > 65536 || mh > 65536) return Error_("baked IBL blob %s: implausible dimensions", filepath);
    memcpy(&ibl->rotationRadians, blob.data() + 48, 4);
    memcpy(&ibl->exposureScale, blob.data() + 52, 4);
    const uint8_t *pm = nullptr, *pc = nullptr, *pl = nullptr, *ps = nullptr;
    ReturnError_(blobPointer(blob, 16, 4 * h, &pm, "marginalDensityFunction"));
    ReturnError_(blobPointer(blob, 24, 4 * w * h, &pc, "conditionalDensityFunctions"));
    ReturnError_(blobPointer(blob, 56, 12 * w * h, &pl, "lightData"));
    ReturnError_(blobPointer(blob, 64, 12 * mw * mh, &ps, "missData"));
    ibl->width = (uint32_t)w; ibl->height = (uint32_t)h; ibl->missWidth = (uint32_t)mw; ibl->missHeight = (uint32_t)mh;
    ibl->marginal.assign((const float*)pm, (const float*)pm + h);
    ibl->conditional.assign((const float*)pc, (const float*)pc + w * h);
    ibl->lightData.assign((const float*)pl, (const float*)pl + 3 * w * h);
    ibl->missData.assign((const float*)ps, (const float*)ps + 3 * mw * mh);
    memset(&ibl->desc, 0, sizeof(ibl->desc));
    ibl->desc.width = ibl->width;
    ibl->desc.height = ibl->height;
    ibl->desc.marginalDensityFunction = ibl->marginal.data();
    ibl->desc.conditionalDensityFunctions = ibl->conditional.data();
    ibl->desc.lightData = ibl->lightData.data();
    ibl->desc.missWidth = ibl->missWidth;
    ibl->desc.missHeight = ibl->missHeight;
    ibl->desc.missData = ibl->missData.data();
    ibl->desc.rotationRadians = ibl->rotationRadians;
    ibl->desc.exposureScale = ibl->exposureScale;
    return Success_;
}

// TextureResourceData as serialized: 0 mipCount u32 | 4 dataSize u32 | 8 mipWidths[16] u32 | 72 mipHeights[16] u32 | 136 mipOffsets[16] u64
//   264 format u32 | 268 pad u32 | 272 texture*                                   (TextureResource.h:15-39, .cpp:23-42)
Error ReadTextureResource(const char* filepath, TextureResource* tex) {
    std::vector<uint8_t> blob;
    ReturnError_(readWholeFile(filepath, blob));
    if (blob.size() < 280) return Error_("baked texture blob %s: %zu bytes is shorter than the 280-byte resource header", filepath, blob.size());
    uint32_t dataSize;
    memcpy(&tex->mipCount, blob.data() + 0, 4);
    memcpy(&dataSize, blob.data() + 4, 4);
    memcpy(tex->mipWidths, blob.data() + 8, 64);
    memcpy(tex->mipHeights, blob.data() + 72, 64);
    memcpy(tex->mipOffsets, blob.data() + 136, 128);
    memcpy(&tex->format, blob.data() + 264, 4);
    if (tex->mipCount == 0 || tex->mipCount > 16 || tex->format > SELAS_TEXTURE_FLOAT4) return Error_("baked texture blob %s: bad mipCount / format", filepath);
    const uint8_t* pt = nullptr;
    ReturnError_(blobPointer(blob, 272, dataSize, &pt, "texture"));
    const uint64_t level0 = (uint64_t)tex->mipWidths[0] * tex->mipHeights[0] * (tex->format + 1) * 4;
    if (tex->mipOffsets[0] > dataSize || level0 > dataSize - tex->mipOffsets[0]) return Error_("baked texture blob %s: level 0 lies outside the data", filepath);
    tex->texture.assign(pt, pt + dataSize);
    tex->desc.width = tex->mipWidths[0];
    tex->desc.height = tex->mipHeights[0];
    tex->desc.format = tex->format;
    tex->desc.reserved = 0;
    tex->desc.texels = reinterpret_cast<const float*>(tex->texture.data() + tex->mipOffsets[0]);
    return Success_;
}

// ---- scene / subscene / model blobs ---------------------------------------------------------------------------------------------
namespace {
std::string fixedString(const uint8_t* p, size_t n) {
    size_t len = 0;
    while (len < n && p[len]) ++len;
    return std::string((const char*)p, len);
}
// CArray header at `fieldOffset`: {flags u32, padding u32, count u64, capacity u64, data offset u64}; returns the element data
Error blobArray(const std::vector<uint8_t>& blob, size_t fieldOffset, size_t elementSize, uint64_t* count, const uint8_t** data, const char* what) {
    if (fieldOffset + 32 > blob.size()) return Error_("baked blob: array header %s lies outside the file", what);
    memcpy(count, blob.data() + fieldOffset + 8, 8);
    if (*count > blob.size()) return Error_("baked blob: array %s has an implausible count", what);
    return blobPointer(blob, fieldOffset + 24, *count * elementSize, data, what);
}
void readMaterial(const uint8_t* p, BakedMaterial& m) {  // MaterialResourceData: 0 texture[512] | 512 shader | 516 flags | 520 baseColor | 532 transmittance | 544 scalars[14]
    m.baseColorTexture = fixedString(p, 512);
    memcpy(&m.material.shader, p + 512, 4);
    memcpy(&m.material.flags, p + 516, 4);
    memcpy(m.material.baseColor, p + 520, 12);
    memcpy(m.material.transmittanceColor, p + 532, 12);
    memcpy(m.material.scalars, p + 544, 4 * SELAS_SCALAR_COUNT);
}
void multiply44(const float* a, const float* b, float* c) {  // MatrixMultiply, MathLib/FloatFuncs.cpp (row-major, row vectors)
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) c[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j] + a[4 * i + 3] * b[12 + j];
}
}  // namespace

// ModelResourceData: 0 aaBox | 24 totalVertexCount | 28 totalCurveVertexCount | 32 curveModelName | 36 pad | 40..103 eight u64 sizes |
//   104 cameras | 136 textureResourceNames | 168 materials | 200 materialHashes | 232 meshes | 264 curves          (296 bytes)
// ModelGeometryData: 0..63 the same eight sizes | 64 indices* | 72 faceIndexCounts* | 80 positions* | 88 normals* | 96 tangents* |
//   104 uvs* | 112 curveIndices* | 120 curveVertices*                                                              (128 bytes)
Error ReadModelResource(const char* dataPath, const char* geometryPath, BakedModel* model) {
    std::vector<uint8_t> blob;
    ReturnError_(readWholeFile(dataPath, blob));
    if (blob.size() < 296) return Error_("baked model blob %s: shorter than the 296-byte resource header", dataPath);
    memcpy(model->aaBox, blob.data(), 24);
    memcpy(&model->totalVertexCount, blob.data() + 24, 4);
    memcpy(&model->totalCurveVertexCount, blob.data() + 28, 4);
    memcpy(&model->curveModelName, blob.data() + 32, 4);
    uint64_t n = 0;
    const uint8_t* p = nullptr;
    ReturnError_(blobArray(blob, 168, 600, &n, &p, "materials"));
    model->materials.resize(n);
    for (uint64_t i = 0; i < n; ++i) readMaterial(p + 600 * i, model->materials[i]);
    ReturnError_(blobArray(blob, 200, 4, &n, &p, "materialHashes"));
    model->materialHashes.assign((const uint32_t*)p, (const uint32_t*)p + n);
    ReturnError_(blobArray(blob, 232, 92, &n, &p, "meshes"));
    model->meshes.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* m = p + 92 * i;
        BakedMesh& bm = model->meshes[i];
        memcpy(&bm.indexCount, m, 4); memcpy(&bm.indexOffset, m + 4, 4); memcpy(&bm.vertexCount, m + 8, 4); memcpy(&bm.vertexOffset, m + 12, 4);
        memcpy(&bm.materialHash, m + 16, 4); memcpy(&bm.indicesPerFace, m + 20, 4); memcpy(&bm.nameHash, m + 24, 4);
        bm.name = fixedString(m + 28, 64);
    }
    ReturnError_(blobArray(blob, 264, 8, &model->curveCount, &p, "curves"));
    model->curveRanges.assign((const uint32_t*)p, (const uint32_t*)p + 2 * model->curveCount);

    std::vector<uint8_t> geo;
    ReturnError_(readWholeFile(geometryPath, geo));
    if (geo.size() < 128) return Error_("baked geometry blob %s: shorter than the 128-byte header", geometryPath);
    uint64_t sizes[8];
    memcpy(sizes, geo.data(), 64);
    const uint8_t *pi = nullptr, *pp = nullptr, *pn = nullptr, *pt = nullptr, *pu = nullptr;
    ReturnError_(blobPointer(geo, 64, sizes[0], &pi, "indices"));
    ReturnError_(blobPointer(geo, 80, sizes[2], &pp, "positions"));
    ReturnError_(blobPointer(geo, 88, sizes[3], &pn, "normals"));
    ReturnError_(blobPointer(geo, 96, sizes[4], &pt, "tangents"));
    ReturnError_(blobPointer(geo, 104, sizes[5], &pu, "uvs"));
    if (sizes[2] != 12ull * model->totalVertexCount) return Error_("baked geometry blob %s: positionSize does not match totalVertexCount", geometryPath);
    model->indices.assign((const uint32_t*)pi, (const uint32_t*)pi + sizes[0] / 4);
    model->positions.assign((const float*)pp, (const float*)pp + sizes[2] / 4);
    model->normals.assign((const float*)pn, (const float*)pn + sizes[3] / 4);
    model->tangents.assign((const float*)pt, (const float*)pt + sizes[4] / 4);
    model->uvs.assign((const float*)pu, (const float*)pu + sizes[5] / 4);
    if (model->curveCount) {
        const uint8_t *pci = nullptr, *pcv = nullptr;
        ReturnError_(blobPointer(geo, 112, sizes[6], &pci, "curveIndices"));
        ReturnError_(blobPointer(geo, 120, sizes[7], &pcv, "curveVertices"));
        if (sizes[7] != 16ull * model->totalCurveVertexCount) return Error_("baked geometry blob %s: curveVertexSize does not match totalCurveVertexCount", geometryPath);
        model->curveIndices.assign((const uint32_t*)pci, (const uint32_t*)pci + sizes[6] / 4);
        model->curveVertices.assign((const float*)pcv, (const float*)pcv + sizes[7] / 4);
    }
    return Success_;
}

// SubsceneResourceData: 0 name[512] | 512 lightSetIndex u64 | 520 modelNames | 552 modelInstances | 584 sceneMaterialNames | 616 sceneMaterials
Error ReadSubsceneResource(const char* filepath, BakedSubscene* sub) {
    std::vector<uint8_t> blob;
    ReturnError_(readWholeFile(filepath, blob));
    if (blob.size() < 648) return Error_("baked subscene blob %s: shorter than the 648-byte resource header", filepath);
    sub->name = fixedString(blob.data(), 512);
    memcpy(&sub->lightSetIndex, blob.data() + 512, 8);
    uint64_t n = 0;
    const uint8_t* p = nullptr;
    ReturnError_(blobArray(blob, 520, 512, &n, &p, "modelNames"));
    sub->modelNames.resize(n);
    for (uint64_t i = 0; i < n; ++i) sub->modelNames[i] = fixedString(p + 512 * i, 512);
    ReturnError_(blobArray(blob, 552, 72, &n, &p, "modelInstances"));
    sub->modelInstances.resize(n);
    for (uint64_t i = 0; i < n; ++i) memcpy(&sub->modelInstances[i], p + 72 * i, 72);
    ReturnError_(blobArray(blob, 584, 4, &n, &p, "sceneMaterialNames"));
    sub->sceneMaterialNames.assign((const uint32_t*)p, (const uint32_t*)p + n);
    ReturnError_(blobArray(blob, 616, 600, &n, &p, "sceneMaterials"));
    sub->sceneMaterials.resize(n);
    for (uint64_t i = 0; i < n; ++i) readMaterial(p + 600 * i, sub->sceneMaterials[i]);
    return Success_;
}

// SceneResourceData: 0 name[512] | 512 iblName[512] | 1024 backgroundIntensity | 1040 subsceneNames | 1072 subsceneInstances | 1104 lights |
//   1136 lightSetRanges | 1168 cameras                                                                            (1200 bytes)
Error ReadSceneResource(const char* filepath, BakedScene* scene) {
    std::vector<uint8_t> blob;
    ReturnError_(readWholeFile(filepath, blob));
    if (blob.size() < 1200) return Error_("baked scene blob %s: shorter than the 1200-byte resource header", filepath);
    scene->name = fixedString(blob.data(), 512);
    scene->iblName = fixedString(blob.data() + 512, 512);
    memcpy(scene->backgroundIntensity, blob.data() + 1024, 16);
    uint64_t n = 0;
    const uint8_t* p = nullptr;
    ReturnError_(blobArray(blob, 1040, 512, &n, &p, "subsceneNames"));
    scene->subsceneNames.resize(n);
    for (uint64_t i = 0; i < n; ++i) scene->subsceneNames[i] = fixedString(p + 512 * i, 512);
    ReturnError_(blobArray(blob, 1072, 72, &n, &p, "subsceneInstances"));
    scene->subsceneInstances.resize(n);
    for (uint64_t i = 0; i < n; ++i) memcpy(&scene->subsceneInstances[i], p + 72 * i, 72);
    ReturnError_(blobArray(blob, 1104, 64, &n, &p, "lights"));
    scene->lights.resize(n);
    if (n) memcpy(scene->lights.data(), p, 64 * n);
    ReturnError_(blobArray(blob, 1136, 8, &n, &p, "lightSetRanges"));
    scene->lightSetRanges.resize(n);
    if (n) memcpy(scene->lightSetRanges.data(), p, 8 * n);
    ReturnError_(blobArray(blob, 1168, 176, &n, &p, "cameras"));
    scene->cameras.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        scene->cameras[i].name = fixedString(p + 176 * i, 128);
        memcpy(&scene->cameras[i].settings, p + 176 * i + 128, 48);  // position, fov, lookAt, znear, up, zfar
    }
    return Success_;
}

Error LoadBakedScene(const char* root, const char* sceneName, SceneResource* scene, ImageBasedLight* ibl, std::vector<BakedCamera>* cameras) {
    BakedScene bs;
    ReturnError_(ReadSceneResource(AssetFilePath(root, "SceneResource", kSceneDataVersion, sceneName).c_str(), &bs));
    scene->meshes.clear();
    scene->curves.clear();
    scene->instances.clear();
    scene->materials.clear();
    scene->flags |= SELAS_SCENE_TWO_LEVEL_BVH;  // a baked scene is instances of models: keep each model's geometry once
    scene->lights = bs.lights;
    scene->lightSets = bs.lightSetRanges;
    memcpy(scene->backgroundIntensity, bs.backgroundIntensity, 16);
    scene->ibl = nullptr;
    if (!bs.iblName.empty()) {
        if (!ibl) return Error_("LoadBakedScene: scene %s names an IBL (%s) but no ImageBasedLight was passed", sceneName, bs.iblName.c_str());
        ReturnError_(ReadImageBasedLightResource(AssetFilePath(root, "IBL", kIblDataVersion, bs.iblName.c_str()).c_str(), ibl));
        scene->ibl = &ibl->desc;
    }
    if (cameras) *cameras = bs.cameras;
    // the default material of CreateDefaultMaterial (ModelResource.cpp:203-211), appended on first use
    int defaultMaterial = -1;
    std::vector<std::string> materialTexture;  // baseColorTexture name of scene->materials[i]
    auto textureFor = [&](const std::string& name, const SelasTextureDesc** out) -> Error {  // TextureCache::LoadTextureResource
        *out = nullptr;
        if (name.empty()) return Success_;
        for (size_t k = 0; k < scene->textureNames.size(); ++k)
            if (scene->textureNames[k] == name) {
                *out = &scene->textures[k]->desc;
                return Success_;
            }
        TextureResource* t = new TextureResource();
        Error e = ReadTextureResource(AssetFilePath(root, "Textures", kTextureDataVersion, name.c_str()).c_str(), t);
        if (e.Failed()) {
            delete t;
            return e;
        }
        scene->textures.push_back(t);
        scene->textureNames.push_back(name);
        *out = &t->desc;
        return Success_;
    };
    // subscenes -> models -> meshes; meshRange[subscene][model] = first global mesh index, mesh count
    struct ModelMeshes { uint32_t first, count, curveFirst, curveCount; };
    std::vector<std::vector<ModelMeshes>> meshRange(bs.subsceneNames.size());
    std::vector<BakedSubscene> subs(bs.subsceneNames.size());
    for (size_t s = 0; s < bs.subsceneNames.size(); ++s) {
        BakedSubscene& sub = subs[s];
        ReturnError_(ReadSubsceneResource(AssetFilePath(root, "SubsceneResource", kSubsceneDataVersion, bs.subsceneNames[s].c_str()).c_str(), &sub));
        const uint32_t sceneMaterialBase = (uint32_t)scene->materials.size();
        for (const BakedMaterial& m : sub.sceneMaterials) {
            scene->materials.push_back(m.material);
            materialTexture.push_back(m.baseColorTexture);
        }
        for (const std::string& modelName : sub.modelNames) {
            BakedModel model;
            ReturnError_(ReadModelResource(AssetFilePath(root, "ModelResource", kModelDataVersion, modelName.c_str()).c_str(),
                                           AssetFilePath(root, "ModelGeometryResource", kModelDataVersion, modelName.c_str()).c_str(), &model));
            const uint32_t modelMaterialBase = (uint32_t)scene->materials.size();
            for (const BakedMaterial& m : model.materials) {
                scene->materials.push_back(m.material);
                materialTexture.push_back(m.baseColorTexture);
            }
            ModelMeshes range{(uint32_t)scene->meshes.size(), (uint32_t)model.meshes.size(), (uint32_t)scene->curves.size(), (uint32_t)model.curveCount};
            // FindMeshMaterial (ModelResource.cpp:214-235): the model's own materials (sorted hashes), then the subscene's, then the default
            auto findMaterial = [&](uint32_t hash) -> int {
                int material = -1;
                auto it = std::lower_bound(model.materialHashes.begin(), model.materialHashes.end(), hash);
                if (it != model.materialHashes.end() && *it == hash) material = (int)(modelMaterialBase + (it - model.materialHashes.begin()));
                for (size_t k = 0; material < 0 && k < sub.sceneMaterialNames.size(); ++k)
                    if (sub.sceneMaterialNames[k] == hash) material = (int)(sceneMaterialBase + k);
                if (material < 0) {
                    if (defaultMaterial < 0) {
                        SelasMaterial dm;
                        memset(&dm, 0, sizeof(dm));
                        dm.shader = SELAS_SHADER_DISNEY_SOLID;
                        dm.baseColor[0] = dm.baseColor[1] = dm.baseColor[2] = 0.6f;
                        dm.scalars[SELAS_IOR] = 1.5f;
                        defaultMaterial = (int)scene->materials.size();
                        scene->materials.push_back(dm);
                        materialTexture.push_back(std::string());
                    }
                    material = defaultMaterial;
                }
                return material;
            };
            // the model's curve geometries (InitializeCurves, ModelResource.cpp:300-333): every curve shares the model-wide control
            // vertices and takes `indexCount` segment indices from `indexOffset`; their material is the one named curveModelName (:382-383)
            for (uint64_t c = 0; c < model.curveCount; ++c) {
                const uint32_t indexOffset = model.curveRanges[2 * c], indexCount = model.curveRanges[2 * c + 1];
                if ((uint64_t)indexOffset + indexCount > model.curveIndices.size())
                    return Error_("LoadBakedScene: curve %llu of model %s indexes outside the curve index buffer", (unsigned long long)c, modelName.c_str());
                SceneCurve curve;
                curve.lightSetIndex = (uint32_t)sub.lightSetIndex;
                curve.materialIndex = (uint32_t)findMaterial(model.curveModelName);
                const uint32_t* ix = model.curveIndices.data() + indexOffset;
                uint32_t lo = 0xFFFFFFFFu, hi = 0;
                for (uint32_t k = 0; k < indexCount; ++k) { lo = std::min(lo, ix[k]); hi = std::max(hi, ix[k] + 3u); }
                if (indexCount == 0) { lo = 0; hi = 0; }
                if (indexCount && hi >= model.totalCurveVertexCount)
                    return Error_("LoadBakedScene: curve %llu of model %s references control vertex %u of %u", (unsigned long long)c, modelName.c_str(), hi,
                                  model.totalCurveVertexCount);
                const uint32_t nv = indexCount ? hi - lo + 1 : 0;
                curve.indices.resize(indexCount);
                for (uint32_t k = 0; k < indexCount; ++k) curve.indices[k] = ix[k] - lo;
                curve.vertices.assign(model.curveVertices.begin() + 4 * (size_t)lo, model.curveVertices.begin() + 4 * (size_t)(lo + nv));
                scene->curves.push_back(std::move(curve));
            }
            for (const BakedMesh& bm : model.meshes) {
                if (bm.indicesPerFace != 3 && bm.indicesPerFace != 4) return Error_("LoadBakedScene: mesh %s has %u indices per face", bm.name.c_str(), bm.indicesPerFace);
                if ((uint64_t)bm.indexOffset + bm.indexCount > model.indices.size()) return Error_("LoadBakedScene: mesh %s indexes outside the model's index buffer", bm.name.c_str());
                SceneMesh mesh;
                mesh.indicesPerFace = bm.indicesPerFace;
                mesh.lightSetIndex = (uint32_t)sub.lightSetIndex;
                const int material = findMaterial(bm.materialHash);
                mesh.materialIndex = (uint32_t)material;
                if (scene->materials[material].flags & SELAS_MAT_USES_PTEX) {
                    // "<material.baseColorTexture>\\<mesh name>.ptx" under the content root (ModelResource.cpp:358-364, AssetFileUtils.cpp:85-92);
                    // the file is opened when the scene goes to the device (InitializeSceneResource), its absence is reported here as
                    // TextureCache::LoadTexturePtex does (TextureCache.cpp:105-108)
                    mesh.ptexPath = std::string(root) + "Content/" + materialTexture[material] + "/" + bm.name + ".ptx";
                    for (char& ch : mesh.ptexPath)
                        if (ch == '\\') ch = '/';
                    FILE* probe = fopen(mesh.ptexPath.c_str(), "rb");
                    if (!probe) return Error_("Ptex file %s does not exist", mesh.ptexPath.c_str());
                    fclose(probe);
                }
                if (mesh.ptexPath.empty()) ReturnError_(textureFor(materialTexture[material], &mesh.baseColorTexture));
                // eDisplacementEnabled materials are plain triangle / quad meshes in the reference as built (EnableDisplacement_ 0,
                // ModelResource.cpp:18,:251-252) and eAlphaTested ones pass every hit (CalculatePassesAlphaTest returns true,
                // SurfaceParameters.cpp:344-347): nothing to do for either
                // the mesh's vertices: indices are absolute into the model-wide arrays (SetMeshVertexAttributes, ModelResource.cpp:164-199);
                // take the span they touch and rebase
                const uint32_t* ix = model.indices.data() + bm.indexOffset;
                uint32_t lo = 0xFFFFFFFFu, hi = 0;
                for (uint32_t k = 0; k < bm.indexCount; ++k) { lo = std::min(lo, ix[k]); hi = std::max(hi, ix[k]); }
                if (bm.indexCount == 0) { lo = 0; hi = 0; }
                if (hi >= model.totalVertexCount) return Error_("LoadBakedScene: mesh %s references vertex %u of %u", bm.name.c_str(), hi, model.totalVertexCount);
                const uint32_t nv = bm.indexCount ? hi - lo + 1 : 0;
                mesh.indices.resize(bm.indexCount);
                for (uint32_t k = 0; k < bm.indexCount; ++k) mesh.indices[k] = ix[k] - lo;
                mesh.positions.assign(model.positions.begin() + 3 * (size_t)lo, model.positions.begin() + 3 * (size_t)(lo + nv));
                if (!model.normals.empty()) mesh.normals.assign(model.normals.begin() + 3 * (size_t)lo, model.normals.begin() + 3 * (size_t)(lo + nv));
                if (!model.tangents.empty()) mesh.tangents.assign(model.tangents.begin() + 4 * (size_t)lo, model.tangents.begin() + 4 * (size_t)(lo + nv));
                if (!model.uvs.empty()) mesh.uvs.assign(model.uvs.begin() + 2 * (size_t)lo, model.uvs.begin() + 2 * (size_t)(lo + nv));
                scene->meshes.push_back(std::move(mesh));
            }
            meshRange[s].push_back(range);
        }
    }
    // flatten the two instancing levels (InitializeModelInstances, SubsceneResource.cpp:63-79; scene level, SceneResource.cpp:218-249)
    // Both levels stay visible in what a hit reports (SELAS_SCENE_INSTANCE_IDS): instID0 = subscene instance, instID1 = model instance
    // inside the subscene, geomID = mesh inside the model -- the triple ModelDataFromRayIds takes (SceneResource.cpp:362-378).
    for (size_t sIdx = 0; sIdx < bs.subsceneInstances.size(); ++sIdx) {
        const BakedInstance& si = bs.subsceneInstances[sIdx];
        if (si.index >= subs.size()) return Error_("LoadBakedScene: subscene instance refers to subscene %llu of %zu", (unsigned long long)si.index, subs.size());
        const BakedSubscene& sub = subs[si.index];
        for (size_t mIdx = 0; mIdx < sub.modelInstances.size(); ++mIdx) {
            const BakedInstance& mi = sub.modelInstances[mIdx];
            if (mi.index >= meshRange[si.index].size()) return Error_("LoadBakedScene: model instance refers to model %llu of %zu", (unsigned long long)mi.index,
                                                                      meshRange[si.index].size());
            const ModelMeshes& range = meshRange[si.index][mi.index];
            SelasInstanceDesc inst;
            multiply44(mi.localToWorld, si.localToWorld, inst.localToWorld);
            inst.instID0 = (uint32_t)sIdx;
            inst.instID1 = (uint32_t)mIdx;
            for (uint32_t m = 0; m < range.count; ++m) {
                inst.meshIndex = range.first + m;
                inst.geomID = m;
                scene->instances.push_back(inst);
            }
            for (uint32_t c = 0; c < range.curveCount; ++c) {  // attached after the meshes (ModelResource.cpp:461-466): geomID = meshes + c
                inst.meshIndex = 0x80000000u | (range.curveFirst + c);  // resolved below, once the number of meshes is final
                inst.geomID = range.count + c;
                scene->instances.push_back(inst);
            }
        }
    }
    for (SelasInstanceDesc& inst : scene->instances)
        if (inst.meshIndex & 0x80000000u) inst.meshIndex = (uint32_t)scene->meshes.size() + (inst.meshIndex & 0x7fffffffu);
    scene->flags |= SELAS_SCENE_INSTANCE_IDS;
    // geometry exists in the scene only through instances: without any, nothing is there to hit (an empty instance list would otherwise
    // mean "every mesh once" to selas_cuda_scene_create)
    if (scene->instances.empty()) {
        scene->meshes.clear();
        scene->curves.clear();
    }
    return Success_;
}

// identical to selas_b200/scenes.py:caustic_scene / sun_sky_ibl
void MakeCausticScene(SceneResource* scene, ImageBasedLight* ibl, uint32_t sphereTheta, uint32_t spherePhi, uint32_t floorCells) {
    const double pi = 3.14159265358979323846;
    scene->meshes.assign(3, SceneMesh());
    for (auto& m : scene->meshes) m.indicesPerFace = 4;
    SceneMesh& floor = scene->meshes[0];
    const uint32_t g = floorCells;
    for (uint32_t i = 0; i <= g; ++i)
        for (uint32_t j = 0; j <= g; ++j) {
            floor.positions.push_back((float)(-6.0 + 12.0 * i / g));
            floor.positions.push_back(0.0f);
            floor.positions.push_back((float)(-6.0 + 12.0 * j / g));
        }
    for (uint32_t i = 0; i < g; ++i)
        for (uint32_t j = 0; j < g; ++j) {
            uint32_t v00 = i * (g + 1) + j;
            uint32_t q[4] = {v00, v00 + 1, v00 + g + 2, v00 + g + 1};  // CCW seen from +y
            floor.indices.insert(floor.indices.end(), q, q + 4);
        }
    floor.materialIndex = 0;
    SceneMesh& sphere = scene->meshes[1];
    const double cx = 0.0, cy = 1.2, cz = 0.0, R = 1.2;
    for (uint32_t t = 0; t <= sphereTheta; ++t)
        for (uint32_t p = 0; p < spherePhi; ++p) {
            double th = pi * t / sphereTheta, ph = 2.0 * pi * p / spherePhi;
            double nx = sin(th) * cos(ph), ny = cos(th), nz = sin(th) * sin(ph);
            sphere.positions.push_back((float)(cx + R * nx));
            sphere.positions.push_back((float)(cy + R * ny));
            sphere.positions.push_back((float)(cz + R * nz));
            sphere.normals.push_back((float)nx);
            sphere.normals.push_back((float)ny);
            sphere.normals.push_back((float)nz);
        }
    for (uint32_t t = 0; t < sphereTheta; ++t)
        for (uint32_t p = 0; p < spherePhi; ++p) {
            uint32_t pn = (p + 1) % spherePhi;
            uint32_t q[4] = {t * spherePhi + p, t * spherePhi + pn, (t + 1) * spherePhi + pn, (t + 1) * spherePhi + p};
            sphere.indices.insert(sphere.indices.end(), q, q + 4);
        }
    sphere.materialIndex = 1;
    SceneMesh& box = scene->meshes[2];
    {
        const float x0 = 2.2f, y0 = 0.0f, z0 = -1.0f, x1 = 3.6f, y1 = 1.6f, z1 = 0.6f;
        float v[8][3] = {{x0, y0, z0}, {x1, y0, z0}, {x1, y0, z1}, {x0, y0, z1}, {x0, y1, z0}, {x1, y1, z0}, {x1, y1, z1}, {x0, y1, z1}};
        addQuad(box, v[4], v[7], v[6], v[5]);
        addQuad(box, v[0], v[4], v[5], v[1]);
        addQuad(box, v[1], v[5], v[6], v[2]);
        addQuad(box, v[2], v[6], v[7], v[3]);
        addQuad(box, v[3], v[7], v[4], v[0]);
    }
    box.materialIndex = 2;
    const char* rough = "{\"shaderName\":\"DisneySolid\",\"baseColor\":[%s],\"ior\":1.2,\"metallic\":0.0,\"specularTint\":0.0,\"roughness\":0.8,"
                        "\"anisotropic\":0.0,\"sheen\":0.0,\"sheenTint\":0.0,\"clearcoat\":0.0,\"clearcoatGloss\":0.0,\"specTrans\":0.0,"
                        "\"diffTrans\":0.0,\"flatness\":0.5}";
    const char* glass = "{\"shaderName\":\"DisneySolid\",\"baseColor\":[0.9,1.0,0.9],\"transmittanceColor\":[0.4519,1.0,0.7054],\"ior\":1.7,"
                        "\"metallic\":0.0,\"specularTint\":0.0,\"roughness\":0.01,\"anisotropic\":0.0,\"sheen\":0.0,\"sheenTint\":0.0,"
                        "\"clearcoat\":0.0,\"clearcoatGloss\":0.0,\"specTrans\":1.0,\"diffTrans\":0.0,\"flatness\":0.0,\"scatterDistance\":15.0}";
    scene->materials.assign(3, SelasMaterial());
    char buf[1024];
    ImportedMaterialData imp;
    snprintf(buf, sizeof(buf), rough, "0.8,0.8,0.78");
    ImportMaterialFromString(buf, &imp);
    BuildMaterial(imp, scene->materials[0]);
    ImportMaterialFromString(glass, &imp);
    BuildMaterial(imp, scene->materials[1]);
    snprintf(buf, sizeof(buf), rough, "0.82,0.39,0.37");
    ImportMaterialFromString(buf, &imp);
    BuildMaterial(imp, scene->materials[2]);
    scene->lights.clear();
    scene->lightSets.clear();
    memset(scene->backgroundIntensity, 0, sizeof(scene->backgroundIntensity));
    // sun_sky_ibl: 0.2 sky with a 16x16-texel 5000-nit sun at theta = 40 degrees; a dimmer quarter-size miss image
    ibl->width = 1024;
    ibl->height = 512;
    ibl->lightData.assign((size_t)3 * ibl->width * ibl->height, 0.2f);
    const uint32_t y0 = (uint32_t)(40.0 / 180.0 * ibl->height) - 8, x0 = ibl->width / 3;
    for (uint32_t y = y0; y < y0 + 16; ++y)
        for (uint32_t x = x0; x < x0 + 16; ++x)
            for (int c = 0; c < 3; ++c) ibl->lightData[3 * ((size_t)y * ibl->width + x) + c] = 5000.0f;
    ibl->missWidth = ibl->width / 4;
    ibl->missHeight = ibl->height / 4;
    ibl->missData.assign((size_t)3 * ibl->missWidth * ibl->missHeight, 0.1f);
    ibl->rotationRadians = 0.3f;
    ibl->exposureScale = 1.0f;
    ibl->BuildDensityFunctions();
    scene->ibl = &ibl->desc;
}

void SetupCausticCamera(uint32_t width, uint32_t height, RayCastCameraSettings& camera) {
    CameraSettings s;
    s.position[0] = 0.0f; s.position[1] = 4.5f; s.position[2] = -9.0f;
    s.lookAt[0] = 0.4f; s.lookAt[1] = 0.6f; s.lookAt[2] = 0.0f;
    s.up[0] = 0.0f; s.up[1] = 1.0f; s.up[2] = 0.0f;
    s.fovDegrees = 45.0f;
    s.znear = 0.1f;
    s.zfar = 10000.0f;
    InitializeRayCastCamera(s, width, height, camera);
}

namespace VCM {
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName, const VCMSettings& settings,
                    SelasVcmStats* stats, Framebuffer* keep) {
    if (!scene->gpuScene) return Error_("VCM::GenerateImage: scene is not initialised on a GPU (there is no CPU fallback)");
    if (settings.iterations == 0) return Error_("VCM::GenerateImage: iterations must be > 0");
    const uint32_t width = (uint32_t)camera.width, height = (uint32_t)camera.height;
    Framebuffer frame;
    FrameBuffer_Initialize(&frame, width, height, 1);  // VCM.cpp:741
    SelasVcmSettings vs;
    memset(&vs, 0, sizeof(vs));
    vs.width = width;
    vs.height = height;
    vs.iterations = settings.iterations;
    vs.maxPathLength = settings.maxBounceCount;
    vs.seed = settings.seed;
    vs.radiusFactor = settings.vcmRadiusFactor;  // sharedData.vcmRadius = VcmRadiusFactor_ * scene->boundingSphere.w, VCM.cpp:778
    vs.radiusAlpha = settings.vcmRadiusAlpha;
    memcpy(vs.boundingSphere, scene->boundingSphere, sizeof(vs.boundingSphere));
    if (selas_cuda_render_vcm(scene->gpuScene, &camera, &vs, &frame.buffers[0][0].x, stats) != SELAS_OK)
        return Error_("%s", selas_cuda_last_error(scene->gpu));
    FrameBuffer_Scale(&frame, 1.0f / (float)settings.iterations);  // FrameBuffer_Normalize(&frame, 1 / iterationsPerPixel), VCM.cpp:801
    ReturnError_(FrameBuffer_Save(&frame, imageName));
    if (keep) *keep = frame;
    FrameBuffer_Shutdown(&frame);
    return Success_;
}
}  // namespace VCM

}  // namespace Selas

extern "C" {
int selas_host_import_material_json(const char* json, SelasMaterial* out, char* errorOut, int errorCapacity) {
    Selas::ImportedMaterialData imp;
    Selas::Error e = Selas::ImportMaterialFromString(json, &imp);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    Selas::BuildMaterial(imp, *out);
    return SELAS_OK;
}
int selas_host_load_baked_scene(const char* root, const char* sceneName, uint64_t* counts8, float* instanceMatrices, uint32_t* instanceMesh,
                                void* materials, char* errorOut, int errorCapacity) {
    Selas::SceneResource scene;
    Selas::ImageBasedLight ibl;
    std::vector<Selas::BakedCamera> cameras;
    Selas::Error e = Selas::LoadBakedScene(root, sceneName, &scene, &ibl, &cameras);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    if (counts8) {
        uint64_t tris = 0, texturedMeshes = 0;
        for (const Selas::SceneMesh& m : scene.meshes) texturedMeshes += m.baseColorTexture ? 1 : 0;
        for (const SelasInstanceDesc& in : scene.instances) {
            if (in.meshIndex >= scene.meshes.size()) {  // a curve geometry: one primitive per segment
                tris += scene.curves[in.meshIndex - scene.meshes.size()].indices.size();
                continue;
            }
            const Selas::SceneMesh& m = scene.meshes[in.meshIndex];
            tris += m.indices.size() / m.indicesPerFace * (m.indicesPerFace == 4 ? 2 : 1);
        }
        uint64_t c[8] = {scene.meshes.size(), scene.instances.size(), scene.materials.size(), scene.lights.size(), scene.lightSets.size(),
                         cameras.size(), tris, (scene.ibl ? 1ull : 0ull) | ((uint64_t)scene.textures.size() << 8) | (texturedMeshes << 32)};
        memcpy(counts8, c, sizeof(c));
    }
    for (size_t i = 0; i < scene.instances.size(); ++i) {
        if (instanceMatrices) memcpy(instanceMatrices + 16 * i, scene.instances[i].localToWorld, 64);
        if (instanceMesh) instanceMesh[i] = scene.instances[i].meshIndex;
    }
    if (counts8) counts8[7] |= (uint64_t)((scene.flags & SELAS_SCENE_INSTANCE_IDS) ? 1 : 0) << 62;
    if (materials && !scene.materials.empty()) memcpy(materials, scene.materials.data(), scene.materials.size() * sizeof(SelasMaterial));
    Selas::ShutdownSceneResource(&scene);
    return SELAS_OK;
}
int selas_host_baked_scene_instance_ids(const char* root, const char* sceneName, uint32_t* ids3, uint64_t capacity, char* errorOut, int errorCapacity) {
    Selas::SceneResource scene;
    Selas::ImageBasedLight ibl;
    Selas::Error e = Selas::LoadBakedScene(root, sceneName, &scene, &ibl, nullptr);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    for (size_t i = 0; i < scene.instances.size() && i < capacity; ++i) {
        ids3[3 * i] = scene.instances[i].instID0;
        ids3[3 * i + 1] = scene.instances[i].instID1;
        ids3[3 * i + 2] = scene.instances[i].geomID;
    }
    Selas::ShutdownSceneResource(&scene);
    return SELAS_OK;
}
int selas_host_baked_scene_bounds(const char* root, const char* sceneName, float* aabb6, float* sphere4, char* errorOut, int errorCapacity) {
    Selas::SceneResource scene;
    Selas::ImageBasedLight ibl;
    Selas::Error e = Selas::LoadBakedScene(root, sceneName, &scene, &ibl, nullptr);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    Selas::CalculateSceneBoundingBox(&scene);
    if (aabb6) memcpy(aabb6, scene.aaBox, sizeof(scene.aaBox));
    if (sphere4) memcpy(sphere4, scene.boundingSphere, sizeof(scene.boundingSphere));
    Selas::ShutdownSceneResource(&scene);
    return SELAS_OK;
}
int selas_host_render_baked_scene(const char* root, const char* sceneName, int device, const SelasCamera* camera, const SelasRenderSettings* settings,
                                  float* film, char* errorOut, int errorCapacity) {
    Selas::SceneResource scene;
    Selas::ImageBasedLight ibl;
    Selas::Error e = Selas::LoadBakedScene(root, sceneName, &scene, &ibl, nullptr);
    if (!e.Failed()) e = Selas::InitializeSceneResource(&scene, device);
    if (!e.Failed()) {
        const size_t plane = (size_t)settings->width * settings->height * 3;
        std::vector<float*> layers(settings->layerCount);
        for (uint32_t l = 0; l < settings->layerCount; ++l) layers[l] = film + plane * l;
        if (selas_cuda_render_pathtracer(scene.gpuScene, camera, settings, layers.data(), nullptr) != SELAS_OK)
            e = Selas::MakeError("%s", selas_cuda_last_error(scene.gpu));
        Selas::ShutdownSceneResource(&scene);
    }
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    return SELAS_OK;
}
int selas_host_read_ibl_blob(const char* path, uint32_t* info4, float* params2, float* marginal, float* conditional, float* light, float* miss,
                             char* errorOut, int errorCapacity) {
    Selas::ImageBasedLight ibl;
    Selas::Error e = Selas::ReadImageBasedLightResource(path, &ibl);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    if (info4) { info4[0] = ibl.width; info4[1] = ibl.height; info4[2] = ibl.missWidth; info4[3] = ibl.missHeight; }
    if (params2) { params2[0] = ibl.rotationRadians; params2[1] = ibl.exposureScale; }
    if (marginal) memcpy(marginal, ibl.marginal.data(), ibl.marginal.size() * 4);
    if (conditional) memcpy(conditional, ibl.conditional.data(), ibl.conditional.size() * 4);
    if (light) memcpy(light, ibl.lightData.data(), ibl.lightData.size() * 4);
    if (miss) memcpy(miss, ibl.missData.data(), ibl.missData.size() * 4);
    return SELAS_OK;
}
int selas_host_read_texture_blob(const char* path, uint32_t* info4, float* texels0, char* errorOut, int errorCapacity) {
    Selas::TextureResource tex;
    Selas::Error e = Selas::ReadTextureResource(path, &tex);
    if (e.Failed()) {
        if (errorOut && errorCapacity > 0) snprintf(errorOut, errorCapacity, "%s", e.Message());
        return SELAS_ERR_FORMAT;
    }
    if (info4) { info4[0] = tex.mipCount; info4[1] = tex.format; info4[2] = tex.desc.width; info4[3] = tex.desc.height; }
    if (texels0) memcpy(texels0, tex.desc.texels, (size_t)tex.desc.width * tex.desc.height * (tex.format + 1) * 4);
    return SELAS_OK;
}
int selas_host_write_hdr(const char* path, uint32_t width, uint32_t height, const float* rgb) {
    return Selas::WriteRadianceHdr(path, width, height, (const Selas::float3*)rgb).Failed() ? SELAS_ERR_ARGUMENT : SELAS_OK;
}
}
