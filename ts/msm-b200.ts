/**
 * TypeScript side of the B200 backend for mitschabaude/msm-zprize: the same three async functions that
 * `createMsm()` (src/msm-batched-affine.ts:573-587), `msmProjective` (src/parallel.ts:69-87) and
 * `createMsmBasic()` (src/msm-basic.ts:34-43) return, over the N-API addon (napi/msm_b200_addon.c) and
 * thus over the C ABI of libmsm_b200.so (include/msm_b200.h).
 *
 * Wiring (the only change inside the reference): in `Weierstraß.create` / `TwistedEdwards.create`
 * (src/parallel.ts:65,199) pick these functions when `process.env.MSM_BACKEND === "b200"`:
 *
 *     const { msm, msmUnsafe } = process.env.MSM_BACKEND === "b200"
 *       ? createMsmB200(Inputs, CURVE_BLS12_377_G1) : createMsm(Inputs);
 *
 * The drivers (scripts/run-msm-377.ts, run-msm-pallas.ts, run-msm-ed-377.ts, scripts/msm-*.ts) stay as they are:
 * they still pass pointers into the curve's wasm memories and get back `{ result, log }` with `result` a
 * pointer to a projective (Weierstraß) or extended (twisted Edwards) point in Montgomery form, which they
 * normalise themselves (scripts/msm-weierstrass.ts:90-92, scripts/msm-twisted-edwards.ts:87).
 *
 * Not compiled or run in this repository's image (no Node); the tested mirror of this file is
 * msm_zprize_b200/parallel.py.
 */
// @ts-ignore -- built by node-gyp from napi/binding.gyp
import addon from "../napi/build/Release/msm_b200.node";

// include/msm_b200.h
export const CURVE_BLS12_377_G1 = 0, CURVE_PALLAS = 1, CURVE_ED_ON_BLS12_377 = 2, CURVE_BLS12_381_G1 = 3;
export const FORM_AFFINE_GLV = 0, FORM_PROJECTIVE = 1, FORM_TE_EXTENDED = 2;
export const LAYOUT_LIMB29_MONT = 0, LAYOUT_LE_BYTES = 1;

type Timing = Record<string, number>;
type AddonResult = { x: Uint8Array; y: Uint8Array; isZero: boolean; timing: Timing };

function bytesToBigint(bytes: Uint8Array): bigint {
  let x = 0n;
  for (let i = bytes.length - 1; i >= 0; i--) x = (x << 8n) | BigInt(bytes[i]);
  return x;
}

function toLog(timing: Timing, verbose: boolean): any[][] {
  // same shape as createLog's rows (src/msm-common.ts:192-230): ["phase... 1.23ms"]
  if (!verbose) return [];
  return Object.entries(timing)
    .filter(([k]) => k !== "windowBits" && k !== "windows" && k !== "rounds")
    .map(([k, v]) => [`${k}... ${v.toFixed(2)}ms`]);
}

/** Resident-bases bookkeeping shared by the three entry points.  The benchmark drivers reuse `pointPtr` over
 *  many runs with fresh scalars (scripts/msm-weierstrass.ts:19,29-33), so the upload + ingest is worth
 *  skipping -- but a wasm address says nothing about its contents (`using ... atCurrentOffset` resets hand the
 *  same offset out again, and randomPointsFast can rewrite a region in place).  Residency is therefore keyed
 *  on a caller-supplied generation: pass `options.basesGeneration` (any value that changes whenever the
 *  points at `pointPtr` change) to keep the bases resident across calls; WITHOUT it every call uploads the
 *  points again, which is what the reference's msm does (it reads the points on every call). */
function makeRunner(contexts: unknown[], Field: any, Scalar: any) {
  // contexts[0] owns the resident bases, the others borrow them (addon.shareBases): with k contexts a caller that
  // issues several msm() calls without awaiting each (Promise.all) keeps k MSMs in flight on the GPU -- the
  // inversions and the bucket reduction of one run beside the rounds of another (+29 % MSMs per second at 2^18
  // points with two contexts, +45 % with four).  A caller that awaits every call, like the reference's drivers,
  // sees exactly the single-context behaviour.
  let basesKey: string | undefined;
  let basesJob: Promise<unknown> = Promise.resolve(); // uploads are serialised and wait for every running MSM
  const free = [...contexts];
  const waiters: ((ctx: unknown) => void)[] = [];
  const acquire = () => (free.length ? Promise.resolve(free.pop()!) : new Promise<unknown>((r) => waiters.push(r)));
  const release = (ctx: unknown) => (waiters.length ? waiters.shift()!(ctx) : free.push(ctx));
  async function withAll<T>(f: () => Promise<T>): Promise<T> {
    const held = await Promise.all(contexts.map(() => acquire()));
    try {
      return await f();
    } finally {
      held.forEach(release);
    }
  }
  const runOn = async (ctx: unknown, scalarPtr: number, N: number, form: number, c: number) =>
    (await addon.run(ctx, Scalar.memoryBytes, scalarPtr, N, LAYOUT_LIMB29_MONT, form, c)) as AddonResult;
  async function upload(pointPtr: number, N: number) {
    await addon.setBases(contexts[0], Field.memoryBytes, pointPtr, N, LAYOUT_LIMB29_MONT);
    for (const other of contexts.slice(1)) addon.shareBases(other, contexts[0]);
  }
  return async function run(scalarPtr: number, pointPtr: number, N: number, form: number, c: number, generation?: unknown) {
    if (generation === undefined) {
      // no residency promise: the points are read on this call, like the reference does -- upload and MSM are one
      // critical section over all contexts
      const job = basesJob.then(() =>
        withAll(async () => {
          basesKey = undefined;
          await upload(pointPtr, N);
          return runOn(contexts[0], scalarPtr, N, form, c);
        })
      );
      basesJob = job.catch(() => undefined);
      return job;
    }
    const key = `${String(generation)}:${pointPtr}:${N}`;
    for (;;) {
      if (key !== basesKey) {
        const job = basesJob.then(() =>
          withAll(async () => {
            if (key === basesKey) return; // another call uploaded the same set meanwhile
            basesKey = undefined;
            await upload(pointPtr, N);
            basesKey = key;
          })
        );
        basesJob = job.catch(() => undefined);
        await job;
      }
      const ctx = await acquire(); // while a context is held no upload can start (it needs all of them)
      if (key !== basesKey) {
        release(ctx); // another point set was uploaded between the check and the acquisition: again
        continue;
      }
      try {
        return await runOn(ctx, scalarPtr, N, form, c);
      } finally {
        release(ctx);
      }
    }
  };
}

type MsmOptions = { c?: number; basesGeneration?: unknown };

/** Optional: page-lock the two wasm memories of a curve (after they have reached their final size), so that the
 *  uploads run as asynchronous DMA at PCIe speed.  Returns the function that undoes it. */
export function pinCurveMemories(Inputs: any): () => void {
  const views = [Inputs.Field.memoryBytes, Inputs.Scalar.memoryBytes];
  for (const v of views) addon.pinMemory(v);
  return () => views.forEach((v) => addon.unpinMemory(v));
}

/** Weierstraß curves: drop-in for `createMsm(Inputs)` plus `msmProjective`. */
export function createMsmB200(Inputs: any, curveId: number, devices: number | number[] = 0, contexts = 1) {
  const { Field, Scalar, Affine, Projective } = Inputs;
  // several devices: ONE msm() call range-shards the points over them inside the library (one host thread per
  // GPU, NCCL gather of the partial points, sum on the first device) -- the shape of the reference's SPMD call
  // over its thread pool (src/threads/threads.ts:354-359, src/msm-batched-affine.ts:294-322)
  const ctxs = Array.from({ length: Math.max(1, contexts) }, () => addon.createContext(curveId, devices));
  const run = makeRunner(ctxs, Field, Scalar);

  async function call(scalarPtr: number, pointPtr: number, N: number, verbose: boolean, form: number,
                      { c = 0, basesGeneration }: MsmOptions) {
    // allocated before the scope so it survives it, like the reference's `result` (src/msm-batched-affine.ts:87-90)
    const result = Field.global.getPointer(Projective.size);
    const r = await run(scalarPtr, pointPtr, N, form, c, basesGeneration);
    using _ = Field.local.atCurrentOffset;
    const affine = Field.local.getPointer(Affine.size);
    // canonical (x, y) -> Montgomery affine point -> projective with Z = mg1 (src/curve-affine.ts:273-284,
    // src/curve-projective.ts:322-333); the zero point only carries the flag
    Affine.writeBigint(affine, { x: bytesToBigint(r.x), y: bytesToBigint(r.y), isZero: r.isZero });
    Projective.fromAffine(result, affine);
    return { result, log: toLog(r.timing, verbose) };
  }

  const msm = (scalarPtr: number, pointPtr: number, N: number, verbose = false, options: MsmOptions = {}) =>
    call(scalarPtr, pointPtr, N, verbose, FORM_AFFINE_GLV, options);
  const msmProjective = (scalarPtr: number, pointPtr: number, N: number, options: MsmOptions = {}) =>
    call(scalarPtr, pointPtr, N, false, FORM_PROJECTIVE, options);
  // the engine always applies the safe addition rules (batchAddNew, src/curve-affine.ts:376-458), which agree
  // with the unsafe ones wherever those are defined
  return { msm, msmUnsafe: msm, msmProjective, destroy: () => [...ctxs].reverse().forEach((c) => addon.destroy(c)) };
}

/** Twisted Edwards (ed-on-bls12-377): drop-in for `createMsmBasic(Inputs)` (src/msm-basic.ts:34-43). */
export function createMsmBasicB200(Inputs: any, devices: number | number[] = 0, contexts = 1) {
  const { Field, Scalar, Curve } = Inputs;
  const ctxs = Array.from({ length: Math.max(1, contexts) }, () => addon.createContext(CURVE_ED_ON_BLS12_377, devices));
  const run = makeRunner(ctxs, Field, Scalar);

  async function msm(scalarPtr: number, pointPtr: number, N: number, { c = 0, basesGeneration }: MsmOptions = {}) {
    const result = Field.global.getPointer(Curve.size);
    const r = await run(scalarPtr, pointPtr, N, FORM_TE_EXTENDED, c, basesGeneration);
    const x = bytesToBigint(r.x), y = bytesToBigint(r.y);
    // extended coordinates of the affine result: (x, y, 1, x*y)  (src/curve-twisted-edwards.ts:435-447)
    Curve.fromBigint(result, { X: x, Y: y, Z: 1n, T: (x * y) % Field.p });
    return { result, log: [] as any[][] };
  }
  return Object.assign(msm, { destroy: () => [...ctxs].reverse().forEach((c) => addon.destroy(c)) });
}
