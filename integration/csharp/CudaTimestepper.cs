// CudaTimestepper : ITimestepper — DefaultTimestepper.Timestep (BepuPhysics/DefaultTimestepper.cs:L28-43) with the Solve stage on the GPU.
// Not compiled here (no .NET toolchain in the build image). Usage: Simulation.Create(pool, narrowPhaseCallbacks, poseIntegratorCallbacks,
// solveDescription, new CudaTimestepper(gravity, linearDamping, angularDamping)).
using System;
using System.Numerics;
using BepuPhysics;
using BepuUtilities;
using BepuCuda;

public unsafe class CudaTimestepper : ITimestepper, IDisposable
{
    public event TimestepperStageHandler BeforeCollisionDetection;   // ITimestepper.cs:L20
    public event TimestepperStageHandler CollisionsDetected;         // ITimestepper.cs:L25
    IntPtr ctx;
    IntegratorDesc integrator;

    public CudaTimestepper(Vector3 gravity, float linearDamping = 0.03f, float angularDamping = 0.03f, int device = 0, bool strict = false)
    {
        Config cfg = default; cfg.DeviceOrdinal = device; cfg.StrictFp = strict ? 1 : 0; cfg.ExecutionMode = 0;
        IntPtr c; Check(Native.bepucuda_create(&cfg, &c)); ctx = c;
        integrator.Gravity[0] = gravity.X; integrator.Gravity[1] = gravity.Y; integrator.Gravity[2] = gravity.Z;
        integrator.LinearDamping = linearDamping; integrator.AngularDamping = angularDamping;   // DemoPoseIntegratorCallbacks (Demos/DemoCallbacks.cs:L12-105)
    }

    public void Timestep(Simulation simulation, float dt, IThreadDispatcher threadDispatcher = null)
    {
        simulation.Sleep(threadDispatcher);
        simulation.PredictBoundingBoxes(dt, threadDispatcher);
        BeforeCollisionDetection?.Invoke(dt, threadDispatcher);
        simulation.CollisionDetection(dt, threadDispatcher);
        CollisionsDetected?.Invoke(dt, threadDispatcher);
        if (!SolveOnDevice(simulation, dt))
            simulation.Solve(dt, threadDispatcher);          // unsupported constraint type this frame: CPU path, as before
        simulation.IncrementallyOptimizeDataStructures(threadDispatcher);
    }

    bool SolveOnDevice(Simulation simulation, float dt)
    {
        var solver = simulation.Solver;
        ref var bodies = ref simulation.Bodies.ActiveSet;
        // SolveDescription (SolveDescription.cs:L21-38): the scheduler is evaluated host-side (Solver_Solve.cs:L743-751).
        var iterations = stackalloc int[solver.SubstepCount];
        for (int i = 0; i < solver.SubstepCount; ++i)
            iterations[i] = solver.VelocityIterationScheduler == null ? solver.VelocityIterationCount : Math.Max(1, solver.VelocityIterationScheduler(i));
        Check(Native.bepucuda_set_solve_description(ctx, solver.SubstepCount, iterations, solver.FallbackBatchThreshold));
        fixed (IntegratorDesc* d = &integrator) Check(Native.bepucuda_set_integrator(ctx, d));
        Check(Native.bepucuda_upload_bodies(ctx, bodies.DynamicsState.Memory, bodies.Count));                       // BodySet.cs:L33

        ref var set = ref solver.ActiveSet;
        Check(Native.bepucuda_begin_constraints(ctx, Vector<float>.Count, set.Batches.Count));
        for (int b = 0; b < set.Batches.Count; ++b)
        {
            ref var batch = ref set.Batches[b];
            for (int t = 0; t < batch.TypeBatches.Count; ++t)
            {
                ref var tb = ref batch.TypeBatches[t];                                                                 // TypeBatch.cs:L10-27
                int rc = Native.bepucuda_upload_type_batch(ctx, b, t, tb.TypeId, tb.ConstraintCount, tb.BodyReferences.Memory, tb.PrestepData.Memory, tb.AccumulatedImpulses.Memory);
                if (rc == -4) return false;   // BEPUCUDA_ERR_UNSUPPORTED_TYPE
                Check(rc);
            }
        }
        var kinematics = stackalloc int[Math.Max(1, solver.ConstrainedKinematicHandles.Count)];                       // Solver.cs:L68
        for (int i = 0; i < solver.ConstrainedKinematicHandles.Count; ++i)
            kinematics[i] = simulation.Bodies.HandleToLocation[solver.ConstrainedKinematicHandles[i]].Index;
        Check(Native.bepucuda_set_constrained_kinematics(ctx, kinematics, solver.ConstrainedKinematicHandles.Count));
        Check(Native.bepucuda_end_constraints(ctx));

        Check(Native.bepucuda_solve(ctx, dt));                                                                         // Simulation.cs:L278-290
        Check(Native.bepucuda_download_bodies(ctx, bodies.DynamicsState.Memory, bodies.Count));
        Check(Native.bepucuda_download_impulses(ctx));   // narrow phase redistributes them next frame (NarrowPhaseConstraintUpdate.cs:L81-135)
        return true;
    }

    void Check(int rc) { if (rc != 0) throw new InvalidOperationException(System.Runtime.InteropServices.Marshal.PtrToStringAnsi(Native.bepucuda_last_error(ctx))); }
    public void Dispose() { if (ctx != IntPtr.Zero) { Native.bepucuda_destroy(ctx); ctx = IntPtr.Zero; } }
}
