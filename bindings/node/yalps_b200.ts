// bindings/node/yalps_b200.ts -- drop-in `simplex` and the new `solveMany` for YALPS on top of the N-API addon.
// NOT COMPILED HERE (no Node/tsc in the build image; tests/test_bindings.py only checks that it calls what addon.c
// exports).  In the YALPS tree this file sits in src/ and replaces one import in BOTH callers of the seam:
//     import { simplex } from "./simplex.js"   ->   import { simplex } from "./yalps_b200.js"
//         src/YALPS.ts:4          (root LP: identity basis)
//         src/branchAndCut.ts:3   (node LPs: applyCuts' output carries the ROOT's permutation, :46-52)
// tableauModel, applyCuts, mostFractionalVar, solution and all types stay untouched.
import { createRequire } from "node:module"
import type { Model, Options, Solution, SolutionStatus } from "./types.js"
import { Tableau, tableauModel, TableauModel } from "./tableau.js"
import { defaultOptions, solution } from "./YALPS.js" // `solution` (src/YALPS.ts:8) needs an `export`

const native = createRequire(import.meta.url)("../build/Release/yalps_b200.node")
const STATUS: SolutionStatus[] = ["optimal", "infeasible", "unbounded", "timedout", "cycled"]

// Created on first use; both throw without a CUDA device: there is no CPU fallback.
let ctx: unknown
let multi: unknown
const getCtx = () => (ctx ??= native.create(0))
const getMulti = (devices?: number[]) => (multi ??= native.createMulti(Int32Array.from(devices ?? [0])))

const packOptions = (o: Required<Options>) =>
  Float64Array.of(o.precision, o.maxPivots, o.tolerance, o.timeout, o.maxIterations, o.checkCycles ? 1 : 0)

// Same contract as src/simplex.ts:144: the WHOLE tableau is mutated in place -- every cell of `matrix` (applyCuts
// builds the cut rows of the node LPs from the final root rows, src/branchAndCut.ts:38-42), `positionOfVariable` and
// `variableAtPosition` -- and the incoming permutation arrays are honoured (yalps_solve_batch_basis).
// `matrix` / the permutations may be subarray views (src/branchAndCut.ts:54-60): N-API hands over the view's own
// start address and length.
export const simplex = (tableau: Tableau, options: Required<Options>): [SolutionStatus, number] => {
  const { width, height, matrix, positionOfVariable, variableAtPosition } = tableau
  const out = new Float64Array(2)
  const rc = native.simplex(getCtx(), height, width, matrix, positionOfVariable, variableAtPosition,
                            packOptions(options), out)
  if (rc !== 0) throw new Error(native.lastError(getCtx()))
  return [STATUS[out[0]], out[1]]
}

// simplex() for ONE fresh tableau too large (or too slow) for one GPU: its rows are dealt over `devices`, one persistent
// kernel per GPU, the pivot row and column travel through NVLink peer memory inside the kernels
// (yalps_multi_solve_large).  Same in-place contract; the tableau must carry the identity permutation (src/tableau.ts:95-98).
export const simplexLarge = (tableau: Tableau, options: Required<Options>, devices: number[]): [SolutionStatus, number] => {
  const { width, height, matrix, positionOfVariable, variableAtPosition } = tableau
  const out = new Float64Array(2)
  const m = getMulti(devices)
  const rc = native.simplexLarge(m, height, width, matrix, positionOfVariable, variableAtPosition, packOptions(options), out)
  if (rc !== 0) throw new Error(native.lastError(m))
  return [STATUS[out[0]], out[1]]
}

// solve()'s numeric part (src/YALPS.ts:77-91) for a model whose tableau would be large and almost all zeros (Vendor
// Selection: 22.6 MB for 9,801 stores).  `stores` is what tableauModel would have written with update(tableau, row, col,
// value) (src/tableau.ts:100-134), in order, as cells[i] = row*width + col and values[i]; a maintainer gets them by
// letting tableauModel push into two arrays instead of allocating `matrix` (later stores to a cell win, as there).
// The device zeroes the matrix and scatters the stores (yalps_solve_sparse); results equal solve() on the dense tableau.
export const solveStores = (
  width: number, height: number, cells: Int32Array, values: Float64Array, integers: readonly number[], sign: number,
  options: Required<Options>,
): { status: SolutionStatus; result: number; height: number; rhs: Float64Array; pos: Int32Array; vars: Int32Array } => {
  const k = integers.length
  const status = new Int32Array(3), result = new Float64Array(2)
  const rhs = new Float64Array(height + 2 * k)
  const pos = new Int32Array(width + height + 2 * k), vars = new Int32Array(width + height + 2 * k)
  const rc = native.solveSparse(getCtx(), height, width, cells, values, Int32Array.from(integers), sign,
                                packOptions(options), status, result, rhs, pos, vars, undefined)
  if (rc !== 0) throw new Error(native.lastError(getCtx()))
  return { status: STATUS[status[0]], result: result[0], height: status[1], rhs, pos, vars }
}

// New API: many models, ONE native call.  All root LPs run as one ragged device batch (sharded over `devices` when
// several GPUs are given); models with integer variables whose root is optimal and fractional then run branch and
// cut on the device side of the ABI, many searches concurrently (yalps_multi_solve_many).  Each result equals
// solve(model, options).
export const solveMany = <VarKey = string, ConKey = string>(
  models: readonly Model<VarKey, ConKey>[],
  options?: Options,
  devices?: number[],
): Solution<VarKey>[] => {
  const opt = { ...defaultOptions, ...options } as Required<Options>
  const tabmods: TableauModel<VarKey, ConKey>[] = models.map(m => tableauModel(m))
  const n = tabmods.length
  if (n === 0) return []

  const heights = new Int32Array(n), widths = new Int32Array(n), signs = new Float64Array(n)
  const offsets = new BigInt64Array(n), intsOffsets = new BigInt64Array(n + 1)
  let cells = 0, nints = 0, rows = 0, perms = 0
  tabmods.forEach(({ tableau: t, integers, sign }, i) => {
    heights[i] = t.height
    widths[i] = t.width
    signs[i] = sign
    offsets[i] = BigInt(cells)
    intsOffsets[i] = BigInt(nints)
    cells += t.height * t.width
    nints += integers.length
    rows += t.height + 2 * integers.length // branch and cut appends up to 2*|integers| cut rows (src/branchAndCut.ts:108)
    perms += t.width + t.height + 2 * integers.length
  })
  intsOffsets[n] = BigInt(nints)
  const matrices = new Float64Array(cells), ints = new Int32Array(Math.max(nints, 1))
  tabmods.forEach(({ tableau: t, integers }, i) => {
    matrices.set(t.matrix, Number(offsets[i]))
    ints.set(integers, Number(intsOffsets[i]))
  })

  const status = new Int32Array(n), result = new Float64Array(n), outHeight = new Int32Array(n)
  const rhs = new Float64Array(rows), pos = new Int32Array(perms), vars = new Int32Array(perms)
  const m = getMulti(devices)
  const rc = native.solveMany(m, heights, widths, offsets, matrices, intsOffsets, ints, signs, packOptions(opt),
                              status, result, outHeight, rhs, pos, vars)
  if (rc !== 0) throw new Error(native.lastError(m))

  // solution() reads column 0 and the two permutation arrays of the final (best) tableau and nothing else
  // (src/YALPS.ts:16-19,32): rebuild exactly that view per model.
  let r0 = 0, p0 = 0
  return tabmods.map((tabmod, i) => {
    const width = widths[i], height = outHeight[i], k = tabmod.integers.length
    const matrix = new Float64Array(height * width)
    for (let r = 0; r < height; r++) matrix[r * width] = rhs[r0 + r]
    const final: Tableau = {
      matrix, width, height,
      positionOfVariable: pos.subarray(p0, p0 + width + height),
      variableAtPosition: vars.subarray(p0, p0 + width + height),
    }
    r0 += heights[i] + 2 * k
    p0 += widths[i] + heights[i] + 2 * k
    return solution({ ...tabmod, tableau: final }, STATUS[status[i]], result[i], opt)
  })
}
