// bindings/node/yalps_b200.ts -- drop-in `simplex` / `solveMany` for YALPS on top of the N-API addon.
// NOT COMPILED HERE (no Node/tsc in the build image).  In the YALPS tree this file replaces the import
//     import { simplex } from "./simplex.js"          (src/YALPS.ts:4, src/branchAndCut.ts:3)
// and leaves tableauModel / solution / types untouched.
import { createRequire } from "node:module"
import type { Model, Options, Solution, SolutionStatus } from "./types.js"
import { Tableau, tableauModel, TableauModel } from "./tableau.js"

const native = createRequire(import.meta.url)("../build/Release/yalps_b200.node")
const ctx = native.create(0) // throws without a CUDA device: there is no CPU fallback
const STATUS: SolutionStatus[] = ["optimal", "infeasible", "unbounded", "timedout", "cycled"]

const packOptions = (o: Required<Options>) =>
  Float64Array.of(o.precision, o.maxPivots, o.tolerance, o.timeout, o.maxIterations, o.checkCycles ? 1 : 0)

// Same contract as src/simplex.ts:144: mutates the tableau's RHS column and permutation arrays in place.
export const simplex = (tableau: Tableau, options: Required<Options>): [SolutionStatus, number] => {
  const { width, height, matrix, positionOfVariable, variableAtPosition } = tableau
  const status = new Int32Array(1), value = new Float64Array(1), pivots = new BigInt64Array(2)
  const rhs = new Float64Array(height)
  const rc = native.solveBatch(ctx, 1, height, width, matrix, packOptions(options), status, value, pivots, rhs,
                               positionOfVariable, variableAtPosition)
  if (rc !== 0) throw new Error(native.lastError(ctx))
  for (let r = 0; r < height; r++) matrix[r * width] = rhs[r] // solution() reads only column 0 (src/YALPS.ts:18-19)
  return [STATUS[status[0]], value[0]]
}

// New: many models, one device batch for all root LPs (uniform shapes shown; ragged via yalps_solve_ragged).
export const solveMany = <V, C>(models: Model<V, C>[], options?: Options): TableauModel<V, C>[] => {
  return models.map(m => tableauModel(m)) // then pack matrices and call native.solveBatch once, see INTEGRATION.md
}
