// VpeNative.cs — P/Invoke binding of include/vpe.h for Unity / .NET.
// NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no dotnet/mono/csc in the image, SURVEY.md §8c);
// it is the reference-side stub a maintainer of rajabala/Volumetric-Particles-For-Unity would add.
// Layouts are LayoutKind.Sequential mirrors of the C structs; tests/test_abi.py pins the C layout
// (sizeof/offsetof) that these declarations must match.
using System;
using System.Runtime.InteropServices;

namespace MetavoxelEngine.Native
{
    [StructLayout(LayoutKind.Sequential)]
    public struct VpeTransform
    {
        public float px, py, pz;          // Transform.position
        public float qx, qy, qz, qw;      // Transform.rotation
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct VpeConfig
    {
        public int numMetavoxelsX, numMetavoxelsY, numMetavoxelsZ;
        public float mvScale;
        public int numVoxelsInMetavoxel;
        public int numBorderVoxels;
        public int rayMarchSteps;
        public float ambientR, ambientG, ambientB;
        public float displacementScale;
        public int fadeOutParticles;
        public float opacityFactor;
        public int softParticleStepDistance;
        public float lightNear, lightFar;
        public float lightCameraDistance;
        public int binMode;
        public float marchEarlyOutTransmittance;
        public int slabZBegin, slabZEnd;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct VpeParticle                // 28 bytes: the ParticleSystem.Particle fields VPR.cs reads
    {
        public float x, y, z;
        public float size;
        public float rotationDeg;
        public float lifetime;
        public float startLifetime;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct VpeCamera
    {
        public VpeTransform transform;
        public float fovYDegrees;
        public int width, height;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct VpeStats
    {
        public int numParticles, numMetavoxelsCovered;
        public long numParticlePairs, voxelsFilled, raySamples;
        public int zBoundary, fillLaunches, marchLaunches;
        public float fillMs, marchMs;
        public long brickPoolBytes;
        public float fillKernelMs, marchKernelMs;
        public long raySamplesSkipped;     // of raySamples: samples in empty space the march does not fetch
    }

    /// VpeDebugOptions of include/vpe.h: experiment / measurement switches, all zero = production behaviour.
    [StructLayout(LayoutKind.Sequential)]
    public struct VpeDebugOptions
    {
        public int marchKernel, noSkip, noGray, noRowPad, marchBands, marchTileLog2W, linkSpinMs, sweepOverlap, noTmaSweep, profileSlices;
        public int noHeadFused;
        public int reserved0, reserved1, reserved2, reserved3, reserved4;
    }

    /// VpeMarchOptions of include/vpe.h: UNORM8 target (particlesRT is ARGB32, VPR.cs:228), the debug views of
    /// SetRaymarchPassConstants (VPR.cs:744-761), the scene depth of mainSceneRT (VPR.cs:204) as eye-space floats.
    [StructLayout(LayoutKind.Sequential)]
    public struct VpeMarchOptions
    {
        public int targetFormat, debugMode;
        public IntPtr sceneDepth;          // pinned float[sceneHeight * sceneWidth] or IntPtr.Zero
        public int sceneWidth, sceneHeight;
    }

    public static class Vpe
    {
        const string Lib = "vpe_cuda";       // libvpe_cuda.so next to the player / in Assets/Plugins/x86_64

        [DllImport(Lib)] public static extern void vpe_default_config(out VpeConfig cfg);
        [DllImport(Lib)] public static extern int vpe_create(ref VpeConfig cfg, int device, out IntPtr ctx);
        [DllImport(Lib)] public static extern int vpe_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern int vpe_set_config(IntPtr ctx, ref VpeConfig cfg);
        [DllImport(Lib)] public static extern int vpe_set_light(IntPtr ctx, ref VpeTransform light, float[] gridCenter);
        [DllImport(Lib)] public static extern int vpe_set_displacement_cubemap(IntPtr ctx, byte[] r8, int edge);
        [DllImport(Lib)] public static extern int vpe_set_light_depth_map(IntPtr ctx, float[] depth01);
        [DllImport(Lib)] public static extern int vpe_fill(IntPtr ctx, [In] VpeParticle[] particles, int n, ref VpeTransform emitter);
        [DllImport(Lib)] public static extern int vpe_march(IntPtr ctx, ref VpeCamera cam, [Out] float[] rgba, [Out] int[] samples);
        // the rest of the frame (VPR.cs:184,204,210)
        [DllImport(Lib)] public static extern int vpe_render_light_depth_map(IntPtr ctx, [In] float[] trianglesWorld, int numTriangles);
        [DllImport(Lib)] public static extern int vpe_set_march_options(IntPtr ctx, ref VpeMarchOptions options);
        [DllImport(Lib)] public static extern int vpe_composite_scene(IntPtr ctx, [In] float[] particlesRgba, [In, Out] float[] sceneRgba, int numPixels, int targetFormat);
        [DllImport(Lib)] public static extern int vpe_get_stats(IntPtr ctx, out VpeStats stats);
        [DllImport(Lib)] public static extern int vpe_set_debug_options(IntPtr ctx, ref VpeDebugOptions options);
        [DllImport(Lib)] public static extern IntPtr vpe_last_error(IntPtr ctx);
        [DllImport(Lib)] public static extern int vpe_abi_version();

        public static string LastError(IntPtr ctx) { return Marshal.PtrToStringAnsi(vpe_last_error(ctx)); }
    }
}
