import sys; sys.path.insert(0,'/root/repo')
from selas_b200 import capi, scenes
ctx = capi.Context(0)
data = scenes.caustic_scene(); sc = capi.Scene(ctx, data)
w, h = 1920, 1080; cam = scenes.caustic_camera(w, h); sph = scenes.bounding_sphere(data)
for it in range(6):
    film, st = sc.render_vcm(cam, capi.vcm_settings(w, h, 1, sph, seed=1, first_iteration=it))
    print(it, f"{st.renderMs:.2f} light {st.lightPassMs:.2f} grid {st.gridBuildMs:.2f} cam {st.cameraPassMs:.2f} merge {st.mergeMs:.2f} conn {st.connectMs:.2f} lv {st.lightVertices}")
film, st = sc.render_vcm(cam, capi.vcm_settings(w, h, 8, sph, seed=1, first_iteration=1))
print('8x', f"{st.renderMs/8:.2f} light {st.lightPassMs/8:.2f} grid {st.gridBuildMs/8:.2f} cam {st.cameraPassMs/8:.2f} merge {st.mergeMs/8:.2f} conn {st.connectMs/8:.2f}")
