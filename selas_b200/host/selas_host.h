// selas_host.h -- C++ host surface above the C ABI (include/selas_cuda.h), mirroring the reference's own entry points
// for this path: same names, argument meaning and error behaviour, re-based on the C++ standard library instead of
// the reference's platform layer.
//
//   Error / Success_ / Error_                         Source/Core/SystemLib/Error.h:11-58
//   ImportMaterial / BuildMaterial                    Source/Core/BuildCommon/ImportMaterial.cpp:17-90, BuildModel.cpp:41-66
//   Framebuffer, FrameBuffer_* / FramebufferWriter    Source/Core/TextureLib/Framebuffer.h:17-48, Framebuffer.cpp:18-138
//   RayCastCameraSettings / InitializeRayCastCamera   Source/Core/GeometryLib/Camera.h:17-47, Camera.cpp:89-115
//   PathTracer::GenerateImage                         Source/Applications/Selas/Source/PathTracer.h:20-21, PathTracer.cpp:301-350
//   VCM::GenerateImage                                Source/Applications/Selas/Source/VCM.h:18, VCM.cpp:738-807
//   BuildImageBasedLight density functions            Source/Core/BuildCommon/BuildImageBasedLight.cpp:24-105
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/selas_cuda.h"

namespace Selas {

// ---- Error: null message == success (Error.h:11-36) ---------------------------------------------------------------
class Error {
public:
    Error() {}
    explicit Error(const std::string& msg) : message_(msg), failed_(true) {}
    bool Success() const { return !failed_; }
    bool Failed() const { return failed_; }
    const char* Message() const { return failed_ ? message_.c_str() : nullptr; }

private:
    std::string message_;
    bool failed_ = false;
};
#define Success_ ::Selas::Error()
Error MakeError(const char* fmt, ...);
#define Error_(...) ::Selas::MakeError(__VA_ARGS__)
#define ReturnError_(expr)                 \
    do {                                   \
        ::Selas::Error e__ = (expr);       \
        if (e__.Failed()) return e__;      \
    } while (0)

// ---- materials ------------------------------------------------------------------------------------------------------
struct ImportedMaterialData {  // ImportMaterial.h
    std::string shaderName;
    std::string baseColorTexture;
    float baseColor[3];
    float transmittanceColor[3];
    float scalarAttributes[SELAS_SCALAR_COUNT];
    std::string scalarAttributeTextures[SELAS_SCALAR_COUNT];
    bool alphaTested, invertDisplacement, usesPtex;
};
Error ImportMaterial(const char* filepath, ImportedMaterialData* material);
Error ImportMaterialFromString(const char* json, ImportedMaterialData* material);
void BuildMaterial(const ImportedMaterialData& imported, SelasMaterial& material);

// ---- film -----------------------------------------------------------------------------------------------------------
struct float3 {
    float x, y, z;
};
struct Framebuffer {  // Framebuffer.h:17-25
    uint32_t width = 0, height = 0, layerCount = 0;
    std::vector<std::vector<float3>> buffers;  // [layer][y * width + x]
};
void FrameBuffer_Initialize(Framebuffer* frame, uint32_t width, uint32_t height, uint32_t layerCount);
void FrameBuffer_Shutdown(Framebuffer* frame);
void FrameBuffer_Scale(Framebuffer* frame, float scale);
Error FrameBuffer_Save(Framebuffer* frame, const char* name);  // writes _Images/<name>_<layer>.hdr (Framebuffer.cpp:43-59)
Error WriteRadianceHdr(const char* path, uint32_t width, uint32_t height, const float3* pixels);

// ---- camera ---------------------------------------------------------------------------------------------------------
typedef SelasCamera RayCastCameraSettings;  // layout-identical (104 bytes)
struct CameraSettings {                     // Camera.h:17-26
    float position[3];
    float fovDegrees;
    float lookAt[3];
    float znear;
    float up[3];
    float zfar;
};
void InitializeRayCastCamera(const CameraSettings& settings, uint32_t width, uint32_t height, RayCastCameraSettings& camera);

// ---- scene ----------------------------------------------------------------------------------------------------------
// What the reference keeps in SceneResource + the Embree handles, reduced to what the path needs.
struct SceneMesh {
    std::vector<float> positions, normals, tangents, uvs;
    std::vector<uint32_t> indices;
    uint32_t indicesPerFace = 3, materialIndex = 0, lightSetIndex = 0;
    SelasPtex* ptex = nullptr;  // borrowed; or leave null and name the file:
    std::string ptexPath;       // Ptex base colour file, loaded (once per distinct path) by InitializeSceneResource and owned by the scene
    const SelasTextureDesc* baseColorTexture = nullptr;  // borrowed (e.g. from SceneResource::textures)
};
struct SceneCurve {  // one RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE geometry (ModelResource.cpp:300-333)
    std::vector<float> vertices;    // float4 control vertices
    std::vector<uint32_t> indices;  // first control vertex per segment
    uint32_t materialIndex = 0, lightSetIndex = 0;
};
struct SceneResource {
    std::vector<SceneMesh> meshes;
    std::vector<SceneCurve> curves;  // curve geometries; an instance names one as meshes.size() + index
    std::vector<SelasInstanceDesc> instances;
    std::vector<SelasMaterial> materials;
    std::vector<SelasLight> lights;
    std::vector<SelasLightSetRange> lightSets;
    float backgroundIntensity[4] = {0, 0, 0, 0};
    uint32_t flags = 0;  // SELAS_SCENE_*: LoadBakedScene sets SELAS_SCENE_TWO_LEVEL_BVH so instanced models are stored once
    const SelasIblDesc* ibl = nullptr;  // SceneResource::iblResource (borrowed), or null: constant background
    std::vector<struct TextureResource*> textures;  // base colour textures loaded by LoadBakedScene (owned; freed by ShutdownSceneResource)
    std::vector<std::string> textureNames;
    std::vector<SelasPtex*> ownedPtex;        // Ptex textures InitializeSceneResource loaded from SceneMesh::ptexPath
    float aaBox[6] = {0, 0, 0, 0, 0, 0};          // SceneResource::aaBox (min xyz, max xyz), filled by InitializeSceneResource
    float boundingSphere[4] = {0, 0, 0, 0};       // SceneResource::boundingSphere: centre of aaBox, half its diagonal
    // device side
    SelasContext* gpu = nullptr;
    SelasScene* gpuScene = nullptr;
};
void CalculateSceneBoundingBox(SceneResource* scene);
Error InitializeSceneResource(SceneResource* scene, int device);  // context + upload + BVH8 build
void ShutdownSceneResource(SceneResource* scene);
void MakeCornellBoxScene(SceneResource* scene, float copperRoughness);  // BASELINE config 1
void SetupCornellCamera(uint32_t width, uint32_t height, RayCastCameraSettings& camera);

// ---- image based light ----------------------------------------------------------------------------------------------
// Owns the arrays a SelasIblDesc points at.  BuildDensityFunctions is the bake step's CDF construction
// (BuildImageBasedLight.cpp:24-105): intensity = sin(theta) * luma, per-row inclusive CDFs, marginal inclusive CDF of the row sums.
struct ImageBasedLight {
    uint32_t width = 0, height = 0, missWidth = 0, missHeight = 0;
    std::vector<float> lightData, missData, marginal, conditional;
    float rotationRadians = 0.f, exposureScale = 1.f;
    SelasIblDesc desc;
    void BuildDensityFunctions();
};
// ---- baked assets -----------------------------------------------------------------------------------
// The reference's bake step writes every resource with SerializeToBinary (IoLib/BinaryStreamSerializer.h:14-32): the resource-data
// struct with each pointer field replaced by the file offset of its data, followed by the pointed-to data; the run time attaches the
// file in place (AttachToBinary, :34-44; CBinaryAttachSerializer, BinarySerializers.cpp:118-139).  These readers take such a file.
// _Assets/<type>/<version>/<name>.bin  (AssetFileUtils::AssetFilePath, Assets/AssetFileUtils.cpp:128-135)
std::string AssetFilePath(const char* root, const char* type, uint64_t version, const char* name);
static const uint64_t kIblDataVersion = 1538087123ull;      // ImageBasedLightResource::kDataVersion, type "IBL"
static const uint64_t kTextureDataVersion = 1534887410ull;  // TextureResource::kDataVersion, type "Textures"
// ImageBasedLightResourceData blob (Serialize, SceneLib/ImageBasedLightResource.cpp:23-44) -> ImageBasedLight incl. its SelasIblDesc
Error ReadImageBasedLightResource(const char* filepath, ImageBasedLight* ibl);
// TextureResourceData blob (Serialize, TextureLib/TextureResource.cpp:23-42): every mip level is kept, `desc` describes level 0
struct TextureResource {
    uint32_t mipCount = 0, format = 0;
    uint32_t mipWidths[16] = {0}, mipHeights[16] = {0};
    uint64_t mipOffsets[16] = {0};
    std::vector<uint8_t> texture;  // all levels, as in the file
    SelasTextureDesc desc;         // level 0 (what the path samples: ForceNoMips_, SurfaceParameters.cpp:21)
};
Error ReadTextureResource(const char* filepath, TextureResource* texture);

// Scene / subscene / model blobs.  Raw region = the resource-data struct (strings are FixedString<512>, arrays are CArray headers
// {flags u32, pad u32, count u64, capacity u64, data*} whose data is the raw element array, ContainersLib/CArray.h:239-250).
static const uint64_t kSceneDataVersion = 1539147471ull;     // SceneResource::kDataVersion, type "SceneResource"
static const uint64_t kSubsceneDataVersion = 1539551841ull;  // SubsceneResource::kDataVersion, type "SubsceneResource"
static const uint64_t kModelDataVersion = 1539636447ull;     // ModelResource::kDataVersion, types "ModelResource" / "ModelGeometryResource"
struct BakedInstance {  // Instance, SubsceneResource.h:28-38 (72 bytes)
    uint64_t index;
    float localToWorld[16];
};
struct BakedMaterial {  // MaterialResourceData, ModelResource.h:68-88 (600 bytes)
    std::string baseColorTexture;
    SelasMaterial material;
};
struct BakedMesh {  // MeshMetaData, ModelResource.h:113-123 (92 bytes)
    uint32_t indexCount, indexOffset, vertexCount, vertexOffset, materialHash, indicesPerFace, nameHash;
    std::string name;
};
struct BakedModel {  // ModelResourceData + ModelGeometryData, ModelResource.h:125-170
    float aaBox[6];
    uint32_t totalVertexCount = 0, totalCurveVertexCount = 0;
    uint32_t curveModelName = 0;  // material hash of the model's curves (ModelResource.cpp:382-383)
    uint64_t curveCount = 0;
    std::vector<uint32_t> curveRanges;   // CurveMetaData {indexOffset, indexCount} per curve geometry (ModelResource.h:107-111)
    std::vector<uint32_t> curveIndices;  // first control vertex of every segment
    std::vector<float> curveVertices;    // float4 {x, y, z, radius}
    std::vector<BakedMaterial> materials;
    std::vector<uint32_t> materialHashes;
    std::vector<BakedMesh> meshes;
    std::vector<uint32_t> indices;
    std::vector<float> positions, normals, tangents, uvs;
};
struct BakedSubscene {  // SubsceneResourceData, SubsceneResource.h:41-49
    std::string name;
    uint64_t lightSetIndex = 0;
    std::vector<std::string> modelNames;
    std::vector<BakedInstance> modelInstances;
    std::vector<uint32_t> sceneMaterialNames;
    std::vector<BakedMaterial> sceneMaterials;
};
struct BakedCamera {  // CameraSettings, Camera.h:17-26 (176 bytes with its 128-byte name)
    std::string name;
    CameraSettings settings;
};
struct BakedScene {  // SceneResourceData, SceneResource.h:49-59
    std::string name, iblName;
    float backgroundIntensity[4];
    std::vector<std::string> subsceneNames;
    std::vector<BakedInstance> subsceneInstances;
    std::vector<SelasLight> lights;
    std::vector<SelasLightSetRange> lightSetRanges;
    std::vector<BakedCamera> cameras;
};
Error ReadModelResource(const char* dataPath, const char* geometryPath, BakedModel* model);
Error ReadSubsceneResource(const char* filepath, BakedSubscene* subscene);
Error ReadSceneResource(const char* filepath, BakedScene* scene);
// ReadSceneResource + InitializeSceneResource of the reference (SceneResource.cpp:266-330) for `<root>_Assets/...`: the scene blob, its
// subscenes, their models and the IBL.  The two instancing levels are flattened: every (subscene instance, model instance, mesh)
// becomes one SelasInstanceDesc with localToWorld = model.localToWorld * subscene.localToWorld (ModelDataFromRayIds,
// SceneResource.cpp:362-378); materials are resolved like FindMeshMaterial (ModelResource.cpp:214-235).  Curves are reported as
// unsupported; displacement and alpha testing are inert in the reference as built (EnableDisplacement_ 0; CalculatePassesAlphaTest
// returns true) and are therefore loaded as plain meshes.  A Ptex material's file is <root>Content/<baseColorTexture>/<mesh name>.ptx
// (ModelResource.cpp:358-364; SceneMesh::ptexPath, opened by InitializeSceneResource).  Other base colour textures are loaded from their "Textures" blobs
// (TextureCache::LoadTextureResource, TextureCache.cpp:68-96: one copy per name; an empty name means the constant colour).
// `scene` is left ready for InitializeSceneResource(scene, device).
Error LoadBakedScene(const char* root, const char* sceneName, SceneResource* scene, ImageBasedLight* ibl, std::vector<BakedCamera>* cameras);

// BASELINE config 5: floor + glass sphere (Content/Materials/BSDF.json) + diffuse box under a sun-and-sky IBL.
void MakeCausticScene(SceneResource* scene, ImageBasedLight* ibl, uint32_t sphereTheta = 96, uint32_t spherePhi = 192, uint32_t floorCells = 64);
void SetupCausticCamera(uint32_t width, uint32_t height, RayCastCameraSettings& camera);

// ---- integrator -----------------------------------------------------------------------------------------------------
// The reference's compile-time macros (PathTracer.cpp:42-46, :248) as runtime settings; defaults = reference values.
struct PathTracerSettings {
    uint32_t maxBounceCount = 2048;  // MaxBounceCount_
    uint32_t pathsPerPixel = 16;     // PathsPerPixel_
    uint32_t layerCount = 2;         // LayerCount_
    uint32_t rrStartBounce = 8;
    uint32_t seed = 0;
};
namespace PathTracer {
// Renders, scales by 1/pathsPerPixel and writes _Images/<imageName>_<layer>.hdr, like the reference.  `stats` optional.
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName,
                    const PathTracerSettings& settings = PathTracerSettings(), SelasRenderStats* stats = nullptr, Framebuffer* keep = nullptr);
}  // namespace PathTracer

// DeferredPathTracer::GenerateImage (DeferredPathTracer.cpp:421-470): SamplesPerPixelX_ x SamplesPerPixelY_ correlated multi-jittered
// primary rays per pixel, one output layer, the ShadeHitPosition estimator (SelasRenderSettings.integrator = 1).
struct DeferredPathTracerSettings {
    uint32_t samplesPerPixelX = 4, samplesPerPixelY = 4;  // :40-41
    uint32_t maxBounceCount = 2048;                       // safety bound only: the reference terminates paths by Russian roulette
    uint32_t seed = 0;
};
namespace DeferredPathTracer {
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName,
                    const DeferredPathTracerSettings& settings = DeferredPathTracerSettings(), SelasRenderStats* stats = nullptr,
                    Framebuffer* keep = nullptr);
}  // namespace DeferredPathTracer

// The reference's VCM macros (VCM.cpp:28-35) as runtime settings; IntegrationSeconds_ becomes an iteration count.
struct VCMSettings {
    uint32_t maxBounceCount = 10;    // MaxBounceCount_
    uint32_t iterations = 16;
    float vcmRadiusFactor = 0.0025f; // VcmRadiusFactor_
    float vcmRadiusAlpha = 0.75f;    // VcmRadiusAlpha_
    uint32_t seed = 0;
};
namespace VCM {
// Renders `iterations` passes, normalises by 1/iterations (FrameBuffer_Normalize, VCM.cpp:801) and writes _Images/<imageName>_0.hdr.
Error GenerateImage(SceneResource* scene, const RayCastCameraSettings& camera, const char* imageName, const VCMSettings& settings = VCMSettings(),
                    SelasVcmStats* stats = nullptr, Framebuffer* keep = nullptr);
}  // namespace VCM

}  // namespace Selas

extern "C" {
// C handles for tests (ctypes): material JSON -> 88-byte payload; Radiance writer.
int selas_host_import_material_json(const char* json, SelasMaterial* out, char* errorOut, int errorCapacity);
int selas_host_write_hdr(const char* path, uint32_t width, uint32_t height, const float* rgb);
// Baked-asset readers for tests: info = {width, height, missWidth, missHeight}, params = {rotationRadians, exposureScale}; any array may be null.
int selas_host_read_ibl_blob(const char* path, uint32_t* info4, float* params2, float* marginal, float* conditional, float* light, float* miss,
                             char* errorOut, int errorCapacity);
// LoadBakedScene summary for tests: counts8 = {meshes, instances, materials, lights, lightSets, cameras, triangles,
// hasIbl | distinct textures << 8 | textured meshes << 32}; when not null,
// instanceMatrices receives 16 floats per instance, instanceMesh one uint32 per instance, materials 88 bytes per material
int selas_host_load_baked_scene(const char* root, const char* sceneName, uint64_t* counts8, float* instanceMatrices, uint32_t* instanceMesh,
                                void* materials, char* errorOut, int errorCapacity);
// Loads the baked scene, builds it on `device` and renders with PathTracer settings through the C ABI; film = layerCount planes of w*h*3
/* SceneResource::aaBox / boundingSphere of a baked scene (CalculateSceneBoundingBox, SceneResource.cpp:194-203): no device needed. */
/* (instID0, instID1, geomID) per flattened instance: the reference's two instancing levels as hits report them. */
int selas_host_baked_scene_instance_ids(const char* root, const char* sceneName, uint32_t* ids3, uint64_t capacity, char* errorOut, int errorCapacity);
int selas_host_baked_scene_bounds(const char* root, const char* sceneName, float* aabb6, float* sphere4, char* errorOut, int errorCapacity);
int selas_host_render_baked_scene(const char* root, const char* sceneName, int device, const SelasCamera* camera, const SelasRenderSettings* settings,
                                  float* film, char* errorOut, int errorCapacity);
// info = {mipCount, format, width0, height0}; texels0 receives level 0 (width0 * height0 * (format + 1) floats) when not null
int selas_host_read_texture_blob(const char* path, uint32_t* info4, float* texels0, char* errorOut, int errorCapacity);
}
