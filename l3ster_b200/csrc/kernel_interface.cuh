// Device-callable mirror of the reference's user-kernel API (include/l3ster/common/KernelInterface.hpp:13-204).
//
// The parameter structure is the reference's: a kernel is a callable `(const Input&, Result&) -> void` that fills
// `A0, A1..AD` (E x U each) and `rhs` (E x n_rhs) per quadrature point, starting from a zero-initialised Result.
// The same source idiom works on the device:
//
//     auto& [operators, rhs] = out;  auto& [A0, Ax, Ay, Az] = operators;  Ax(0, 1) = -k;  rhs[0] = s;
//     const auto& [vals, ders, point] = in;   in.normal[0];   in.point.space.x();
//
// Differences forced by the device: the callable is a by-value functor whose operator() is `__host__ __device__`
// (reference kernels are host lambdas, some capturing by reference), and Eigen types are replaced by the tiny fixed-size
// types below (same element access syntax, same column-major storage).
#ifndef L3B_KERNEL_INTERFACE_CUH
#define L3B_KERNEL_INTERFACE_CUH

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

#ifdef __CUDACC__
#define L3B_HD __host__ __device__ __forceinline__
#else
#define L3B_HD inline
#endif

namespace l3b
{
using val_t = double;
using dim_t = unsigned char;
using std::size_t;

// common/KernelInterface.hpp:13-20
struct KernelParams
{
    int    dimension;
    size_t n_equations;
    size_t n_unknowns = 1;
    size_t n_fields   = 0;
    size_t n_rhs      = 1;
};

// fixed-size array with the tuple protocol (structured bindings on host and device)
template < typename T, size_t N >
struct Array
{
    T                       v[N > 0 ? N : 1];
    L3B_HD constexpr T&       operator[](size_t i) { return v[i]; }
    L3B_HD constexpr const T& operator[](size_t i) const { return v[i]; }
    L3B_HD static constexpr size_t size() { return N; }
    template < size_t I >
    L3B_HD constexpr T& get()
    {
        static_assert(I < N);
        return v[I];
    }
    template < size_t I >
    L3B_HD constexpr const T& get() const
    {
        static_assert(I < N);
        return v[I];
    }
};

// column-major R x C matrix: the storage order of Eigen::Matrix<val_t, R, C> (KernelInterface.hpp:32-33)
template < size_t R, size_t C >
struct Matrix
{
    val_t                       v[R * C > 0 ? R * C : 1];
    L3B_HD constexpr val_t&       operator()(size_t i, size_t j) { return v[i + j * R]; }
    L3B_HD constexpr const val_t& operator()(size_t i, size_t j) const { return v[i + j * R]; }
    L3B_HD constexpr val_t&       operator[](size_t i) { return v[i]; } // vector-style access, as Eigen allows for C == 1
    L3B_HD constexpr const val_t& operator[](size_t i) const { return v[i]; }
    L3B_HD constexpr void         setZero()
    {
        for (size_t i = 0; i < R * C; ++i)
            v[i] = 0.;
    }
    static constexpr size_t rows = R, cols = C;
};

// common/Structs.hpp:10-86
template < size_t DIM >
struct Point
{
    val_t                       coords[DIM];
    L3B_HD constexpr val_t        operator[](size_t i) const { return coords[i]; }
    L3B_HD constexpr val_t&       operator[](size_t i) { return coords[i]; }
    L3B_HD constexpr val_t        x() const { return coords[0]; }
    L3B_HD constexpr val_t        y() const
    {
        static_assert(DIM >= 2);
        return coords[1];
    }
    L3B_HD constexpr val_t z() const
    {
        static_assert(DIM >= 3);
        return coords[2];
    }
};
struct SpaceTimePoint
{
    Point< 3 > space;
    val_t      time;
};

// common/KernelInterface.hpp:29-57
template < KernelParams params >
struct KernelInterface
{
    using Operator = Matrix< params.n_equations, params.n_unknowns >;
    using Rhs      = Matrix< params.n_equations, params.n_rhs >;
    struct Result
    {
        Array< Operator, params.dimension + 1 > operators;
        Rhs                                     rhs;
    };
    using FieldVals = Array< val_t, params.n_fields >;
    using FieldDers = Array< FieldVals, params.dimension >;
    using Normal    = Array< val_t, params.dimension >;
    struct DomainInput
    {
        FieldVals      field_vals;
        FieldDers      field_ders;
        SpaceTimePoint point;
    };
    struct BoundaryInput
    {
        FieldVals      field_vals;
        FieldDers      field_ders;
        SpaceTimePoint point;
        Normal         normal;
    };
};

// common/KernelInterface.hpp:102-138: wrappers zero-initialise the result, then invoke the user callable
template < typename Kernel, KernelParams params >
struct DomainEquationKernel
{
    static constexpr auto parameters  = params;
    static constexpr bool is_boundary = false;
    using Input                       = typename KernelInterface< params >::DomainInput;
    using Result                      = typename KernelInterface< params >::Result;
    using functor_type                = Kernel;
    constexpr DomainEquationKernel(Kernel kernel) : m_kernel{kernel} {}
    L3B_HD constexpr Result operator()(const Input& input) const
    {
        Result retval{};
        for (size_t i = 0; i <= static_cast< size_t >(params.dimension); ++i)
            retval.operators[i].setZero();
        retval.rhs.setZero();
        m_kernel(input, retval);
        return retval;
    }
    Kernel m_kernel;
};
template < typename Kernel, KernelParams params >
struct BoundaryEquationKernel
{
    static constexpr auto parameters  = params;
    static constexpr bool is_boundary = true;
    using Input                       = typename KernelInterface< params >::BoundaryInput;
    using Result                      = typename KernelInterface< params >::Result;
    using functor_type                = Kernel;
    constexpr BoundaryEquationKernel(Kernel kernel) : m_kernel{kernel} {}
    L3B_HD constexpr Result operator()(const Input& input) const
    {
        Result retval{};
        for (size_t i = 0; i <= static_cast< size_t >(params.dimension); ++i)
            retval.operators[i].setZero();
        retval.rhs.setZero();
        m_kernel(input, retval);
        return retval;
    }
    Kernel m_kernel;
};

// common/KernelInterface.hpp:178-190
template < KernelParams params, typename Kernel >
constexpr auto wrapDomainEquationKernel(Kernel kernel)
{
    return DomainEquationKernel< Kernel, params >{kernel};
}
template < KernelParams params, typename Kernel >
constexpr auto wrapBoundaryEquationKernel(Kernel kernel)
{
    return BoundaryEquationKernel< Kernel, params >{kernel};
}

// common/KernelInterface.hpp:140-176, 192-204: residual kernels `(const Input&, Rhs&) -> void`, the integrands of
// computeIntegral / computeNormL2 (post/Integral.hpp, post/NormL2.hpp). The reference leaves the result uninitialised before the call;
// here it starts from zero.
template < typename Kernel, KernelParams params >
struct ResidualDomainKernel
{
    static constexpr auto parameters  = params;
    static constexpr bool is_boundary = false;
    using Input                       = typename KernelInterface< params >::DomainInput;
    using Rhs                         = typename KernelInterface< params >::Rhs;
    using functor_type                = Kernel;
    constexpr ResidualDomainKernel(Kernel kernel) : m_kernel{kernel} {}
    L3B_HD constexpr Rhs operator()(const Input& input) const
    {
        Rhs retval{};
        retval.setZero();
        m_kernel(input, retval);
        return retval;
    }
    Kernel m_kernel;
};
template < typename Kernel, KernelParams params >
struct ResidualBoundaryKernel
{
    static constexpr auto parameters  = params;
    static constexpr bool is_boundary = true;
    using Input                       = typename KernelInterface< params >::BoundaryInput;
    using Rhs                         = typename KernelInterface< params >::Rhs;
    using functor_type                = Kernel;
    constexpr ResidualBoundaryKernel(Kernel kernel) : m_kernel{kernel} {}
    L3B_HD constexpr Rhs operator()(const Input& input) const
    {
        Rhs retval{};
        retval.setZero();
        m_kernel(input, retval);
        return retval;
    }
    Kernel m_kernel;
};
template < KernelParams params, typename Kernel >
constexpr auto wrapDomainResidualKernel(Kernel kernel)
{
    return ResidualDomainKernel< Kernel, params >{kernel};
}
template < KernelParams params, typename Kernel >
constexpr auto wrapBoundaryResidualKernel(Kernel kernel)
{
    return ResidualBoundaryKernel< Kernel, params >{kernel};
}

// ---------------------------------------------------------------------------------------------------------------------
// Structural sparsity of a kernel's operators, discovered at compile time.
//
// Least-squares kernels fill a handful of the (D+1) x E x U operator entries (15 of 112 for 3-D diffusion) on top of the
// zero-initialised Result. IEEE arithmetic forbids the compiler from dropping `0.0 * x`, so the device kernels skip the
// dead multiply-adds explicitly: the kernel functor is evaluated *at compile time* on two generic probe inputs, entries
// that are exactly zero for both are treated as structurally zero, and every device evaluation re-checks that those
// entries really are zero (a check the compiler folds away for entries it can prove constant) — a violation raises
// L3B_ERR_SPARSITY instead of producing a wrong result. Kernels whose functor is not constexpr-evaluable (e.g. it calls
// a transcendental) simply get the dense mask.
template < typename KernelT >
constexpr auto makeProbeInput(int variant)
{
    typename KernelT::Input in{};
    constexpr auto          params = KernelT::parameters;
    const val_t             s      = variant == 0 ? 1. : -1.37;
    for (size_t f = 0; f < params.n_fields; ++f)
    {
        in.field_vals[f] = s * (0.37 + 0.11 * static_cast< val_t >(f));
        for (size_t d = 0; d < static_cast< size_t >(params.dimension); ++d)
            in.field_ders[d][f] = s * (-0.23 + 0.07 * static_cast< val_t >(d) + 0.13 * static_cast< val_t >(f));
    }
    in.point.space.coords[0] = 0.123 * s;
    in.point.space.coords[1] = 0.456 * s;
    in.point.space.coords[2] = 0.789 * s;
    in.point.time            = 0.31 * s;
    if constexpr (KernelT::is_boundary)
        for (size_t d = 0; d < static_cast< size_t >(params.dimension); ++d)
            in.normal[d] = s * (0.48 + 0.16 * static_cast< val_t >(d));
    return in;
}

template < typename KernelT >
constexpr auto probeOperatorMask()
{
    constexpr auto params = KernelT::parameters;
    constexpr size_t n    = (params.dimension + 1) * params.n_equations * params.n_unknowns;
    Array< bool, n > mask{};
    for (int variant = 0; variant < 2; ++variant)
    {
        const KernelT kernel{typename KernelT::functor_type{}};
        const auto    res = kernel(makeProbeInput< KernelT >(variant));
        for (size_t i = 0; i <= static_cast< size_t >(params.dimension); ++i)
            for (size_t k = 0; k < params.n_equations * params.n_unknowns; ++k)
                mask[i * params.n_equations * params.n_unknowns + k] =
                    mask[i * params.n_equations * params.n_unknowns + k] or res.operators[i].v[k] != 0.;
    }
    return mask;
}

template < typename KernelT >
concept ConstexprProbeable = requires { typename std::integral_constant< size_t, (probeOperatorMask< KernelT >(), size_t{0}) >; };

template < typename KernelT >
struct KernelSparsity
{
    static constexpr auto   params = KernelT::parameters;
    static constexpr size_t E = params.n_equations, U = params.n_unknowns, n = (params.dimension + 1) * E * U;
    static constexpr bool   probed = ConstexprProbeable< KernelT >;
    static constexpr auto   mask   = [] {
        if constexpr (ConstexprProbeable< KernelT >)
            return probeOperatorMask< KernelT >();
        else
        {
            Array< bool, n > m{};
            for (size_t i = 0; i < n; ++i)
                m[i] = true;
            return m;
        }
    }();
    // is operator `op` entry (eq, u) structurally non-zero?
    static constexpr bool nz(size_t op, size_t eq, size_t u) { return mask[op * E * U + eq + u * E]; }
    static constexpr size_t count()
    {
        size_t c = 0;
        for (size_t i = 0; i < n; ++i)
            c += mask[i];
        return c;
    }
};

// compile-time loop: f(std::integral_constant<int, I>{}) for I in [0, N)
template < int N, typename F >
L3B_HD constexpr void staticFor(F&& f)
{
    [&]< int... I >(std::integer_sequence< int, I... >) { (f(std::integral_constant< int, I >{}), ...); }(std::make_integer_sequence< int, N >{});
}

// algsys/AssembleLocalSystem.hpp:16-49
enum struct LocalEvalStrategy : int
{
    Auto                                 = 0,
    LocalElement                         = 1,
    SumFactorization                     = 2,
    SumFactorizationOddEvenDecomposition = 3
};
struct AssemblyOptions
{
    int               value_order      = 1;
    int               derivative_order = 0;
    LocalEvalStrategy eval_strategy    = LocalEvalStrategy::Auto;
    constexpr int     order(int elem_order) const { return value_order * elem_order + derivative_order * (elem_order - 1); }
    // quad/ReferenceQuadrature.hpp:13-22 applied to QO = 2 * order(EO) (algsys/AssembleGlobalSystem.hpp:42)
    constexpr int nq1d(int elem_order) const { return (2 * order(elem_order)) / 2 + 1; }
};
} // namespace l3b

namespace std
{
template < typename T, size_t N >
struct tuple_size< l3b::Array< T, N > > : integral_constant< size_t, N >
{};
template < size_t I, typename T, size_t N >
struct tuple_element< I, l3b::Array< T, N > >
{
    using type = T;
};
} // namespace std
#endif
