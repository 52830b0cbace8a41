// VolumetricParticleRendererNative.cs — how the reference's MonoBehaviour would call the native path.
// NOT COMPILED HERE (no .NET toolchain in this image).  It replaces exactly the hot path of
// Assets/Main Scene/VolumetricParticleRenderer.cs:
//   BinParticlesToMetavoxels + FillMetavoxels   (VPR.cs:397-520)  -> Vpe.vpe_fill
//   RenderMetavoxels                            (VPR.cs:637-713)  -> Vpe.vpe_march
// and keeps everything else of the reference (scene render into mainSceneRT, CompositeParticles
// blit, GUI setters, debug grid) as it is.  The image comes back as float RGBA and is uploaded into
// particlesRT's stand-in texture before the reference's own composite blit (VPR.cs:210).
using System;
using UnityEngine;
using MetavoxelEngine.Native;

namespace MetavoxelEngine
{
    public partial class VolumetricParticleRenderer : MonoBehaviour
    {
        IntPtr vpe = IntPtr.Zero;
        VpeParticle[] nativeParticles;
        ParticleSystem.Particle[] unityParticles;
        float[] rgba;
        Texture2D particlesTex;   // RGBAFloat stand-in for particlesRT (VPR.cs:228)

        VpeConfig NativeConfig()
        {
            VpeConfig c;
            Vpe.vpe_default_config(out c);
            c.numMetavoxelsX = numMetavoxelsX; c.numMetavoxelsY = numMetavoxelsY; c.numMetavoxelsZ = numMetavoxelsZ;
            c.mvScale = mvScale.x;                                   // cubic metavoxels only (VPR.cs:422,425,445)
            c.numVoxelsInMetavoxel = numVoxelsInMetavoxel;
            c.numBorderVoxels = numBorderVoxels;
            c.rayMarchSteps = rayMarchSteps;
            c.ambientR = ambientColor.x; c.ambientG = ambientColor.y; c.ambientB = ambientColor.z;
            c.displacementScale = fDisplacementScale;
            c.fadeOutParticles = fadeOutParticles ? 1 : 0;
            c.opacityFactor = opacityFactor;
            c.softParticleStepDistance = softParticleStepDistance;
            return c;
        }

        static VpeTransform ToNative(Transform t)
        {
            return new VpeTransform { px = t.position.x, py = t.position.y, pz = t.position.z,
                                      qx = t.rotation.x, qy = t.rotation.y, qz = t.rotation.z, qw = t.rotation.w };
        }

        // called from Start() in place of CreateMetavoxelGrid's per-metavoxel RenderTextures (VPR.cs:285-317)
        void StartNative(Cubemap displacement)
        {
            VpeConfig c = NativeConfig();
            if (Vpe.vpe_create(ref c, 0, out vpe) != 0) { Debug.LogError("[VPE] vpe_create failed"); return; }
            int e = displacement.width;
            byte[] r8 = new byte[6 * e * e];
            for (int f = 0; f < 6; f++)
            {
                Color[] px = displacement.GetPixels((CubemapFace)f);  // +X,-X,+Y,-Y,+Z,-Z = the stored face order
                for (int i = 0; i < e * e; i++) r8[f * e * e + i] = (byte)Mathf.RoundToInt(px[i].r * 255f);
            }
            Vpe.vpe_set_displacement_cubemap(vpe, r8, e);
            rgba = new float[Screen.width * Screen.height * 4];
            particlesTex = new Texture2D(Screen.width, Screen.height, TextureFormat.RGBAFloat, false, true);
        }

        // replaces the body of `if (Time.frameCount % updateInterval == 0)` (VPR.cs:186-199)
        void FillNative()
        {
            VpeConfig c = NativeConfig();
            Vpe.vpe_set_config(vpe, ref c);
            VpeTransform light = ToNative(dirLight.transform);
            Vector3 g = gridCenter.transform.position;
            Vpe.vpe_set_light(vpe, ref light, new float[] { g.x, g.y, g.z });
            if (unityParticles == null || unityParticles.Length < particleSys.maxParticles)
            {
                unityParticles = new ParticleSystem.Particle[particleSys.maxParticles];
                nativeParticles = new VpeParticle[particleSys.maxParticles];
            }
            int n = particleSys.GetParticles(unityParticles);        // VPR.cs:412-413
            for (int i = 0; i < n; i++)
            {
                ParticleSystem.Particle p = unityParticles[i];
                nativeParticles[i] = new VpeParticle { x = p.position.x, y = p.position.y, z = p.position.z, size = p.size,
                                                       rotationDeg = p.rotation, lifetime = p.lifetime, startLifetime = p.startLifetime };
            }
            VpeTransform emitter = ToNative(particleSys.transform);
            if (Vpe.vpe_fill(vpe, nativeParticles, n, ref emitter) != 0) Debug.LogError("[VPE] " + Vpe.LastError(vpe));
        }

        // replaces RenderMetavoxels() (VPR.cs:207, 637-713); the result feeds the reference's composite blit (VPR.cs:210)
        void RenderNative()
        {
            Camera cam = Camera.main;
            VpeCamera vc = new VpeCamera { transform = ToNative(cam.transform), fovYDegrees = cam.fieldOfView,
                                           width = Screen.width, height = Screen.height };
            if (Vpe.vpe_march(vpe, ref vc, rgba, null) != 0) { Debug.LogError("[VPE] " + Vpe.LastError(vpe)); return; }
            particlesTex.SetPixelData(rgba, 0);
            particlesTex.Apply(false);
            Graphics.Blit(particlesTex, mainSceneRT, matBlendParticles);
        }

        // replaces `lightCamera.GetComponent<Camera>().RenderWithShader(generateLightDepthMapShader, ...)` (VPR.cs:184):
        // the meshes of the Default layer (the light camera's cullingMask, VPR.cs:346) as world-space triangles,
        // Unity winding as is (GenerateLightDepthMap.shader:6 culls front faces).
        void LightDepthMapNative(MeshFilter[] occluders)
        {
            var tris = new System.Collections.Generic.List<float>();
            foreach (MeshFilter mf in occluders)
            {
                Mesh m = mf.sharedMesh;
                Vector3[] v = m.vertices;
                int[] idx = m.triangles;
                Matrix4x4 l2w = mf.transform.localToWorldMatrix;
                for (int i = 0; i < idx.Length; i++)
                {
                    Vector3 w = l2w.MultiplyPoint3x4(v[idx[i]]);
                    tris.Add(w.x); tris.Add(w.y); tris.Add(w.z);
                }
            }
            if (Vpe.vpe_render_light_depth_map(vpe, tris.ToArray(), tris.Count / 9) != 0) Debug.LogError("[VPE] " + Vpe.LastError(vpe));
        }

        // the debug views of SetRaymarchPassConstants (VPR.cs:744-761) and the 8-bit particlesRT (VPR.cs:228);
        // sceneDepth (eye-space depth of mainSceneRT, VPR.cs:204) is left out here: pass a pinned float[] if wanted
        void MarchOptionsNative(bool eightBitTarget)
        {
            VpeMarchOptions o = new VpeMarchOptions();
            o.targetFormat = eightBitTarget ? 1 : 0;
            o.debugMode = bShowMetavoxelDrawOrder ? 1 : bShowRayMarchBlendFunc ? 2 : bShowRayMarchSamplesPerPixel ? 3 : 0;
            o.sceneDepth = IntPtr.Zero;
            Vpe.vpe_set_march_options(vpe, ref o);
        }

        void OnDestroy() { if (vpe != IntPtr.Zero) Vpe.vpe_destroy(vpe); }
    }
}
