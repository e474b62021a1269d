// formats.cpp -- storage classes of the host layer (DenseMatrix, SparseMatrix{CSR,COO,ELL,BSR}).
// Behavioural mirror of the reference's src/formats/*.cu (text parsers, pinned-host / device
// allocation with zero fill, H2D copies, toDense, cuSPARSE descriptors); conversions and the
// dense transpose run on the device through the C ABI (include/cuspmm_b200.h).
#include "format.hpp"

#include <limits>
#include <vector>

namespace cuspmm {

static std::ifstream openOrThrow(const std::string &path) {
    std::ifstream f(path);
    if (!f.is_open()) {
        std::cerr << "File " << path << " doesn't exist!" << std::endl;
        throw std::runtime_error("cannot open " + path);
    }
    return f;
}

template <typename DT> static cudaDataType cudaTypeOf() {
    if constexpr (std::is_same_v<DT, float>) return CUDA_R_32F;
    else return CUDA_R_64F;
}

static cudaMemcpyKind kindOf(bool srcDev, bool dstDev) {
    if (srcDev && dstDev) return cudaMemcpyDeviceToDevice;
    if (srcDev) return cudaMemcpyDeviceToHost;
    if (dstDev) return cudaMemcpyHostToDevice;
    return cudaMemcpyHostToHost;
}

// =============================================================================== DenseMatrix
// dense.in: "rows cols [nnz]" then one text row per line (reference reader: src/formats/dense.cu:9-36)
template <typename DT, typename MT>
DenseMatrix<DT, MT>::DenseMatrix(std::string filePath) {
    this->ordering = ORDERING::ROW_MAJOR;
    std::ifstream in = openOrThrow(filePath);
    std::string line;
    in >> this->numRows >> this->numCols;
    std::getline(in, line);
    allocateSpace(false);
    for (MT r = 0; r < this->numRows; ++r) {
        std::getline(in, line);
        std::istringstream row(line);
        for (MT c = 0; c < this->numCols; ++c) row >> this->data[RowMjIdx(r, c, this->numCols)];
    }
}

template <typename DT, typename MT>
DenseMatrix<DT, MT>::DenseMatrix(MT numRows, MT numCols, bool onDevice, ORDERING ordering) {
    this->numRows = numRows;
    this->numCols = numCols;
    this->ordering = ordering;
    allocateSpace(onDevice);
}

template <typename DT, typename MT>
DenseMatrix<DT, MT>::DenseMatrix(DenseMatrix<DT, MT> *source, bool onDevice) {
    this->numRows = source->numRows;
    this->numCols = source->numCols;
    this->ordering = source->ordering;
    allocateSpace(onDevice);
    copyData(source);
}

template <typename DT, typename MT>
bool DenseMatrix<DT, MT>::copyData(DenseMatrix<DT, MT> *source) {
    assertSameShape(source);
    cudaCheckError(cudaMemcpy(this->data, source->data, numElements() * sizeof(DT), kindOf(source->onDevice, this->onDevice)));
    return true;
}

template <typename DT, typename MT>
void DenseMatrix<DT, MT>::setCusparseDnMatDesc(cusparseDnMatDescr_t *matDescP) {
    const bool rowMajor = this->ordering == ORDERING::ROW_MAJOR;
    CHECK_CUSPARSE(cusparseCreateDnMat(matDescP, this->numRows, this->numCols, rowMajor ? this->numCols : this->numRows,
                                       this->data, cudaTypeOf<DT>(), rowMajor ? CUSPARSE_ORDER_ROW : CUSPARSE_ORDER_COL));
}

template <typename DT, typename MT>
void DenseMatrix<DT, MT>::assertSameShape(DenseMatrix<DT, MT> *target) {
    assert(this->numRows == target->numRows && this->numCols == target->numCols);
    if (this->numRows != target->numRows || this->numCols != target->numCols) throw std::runtime_error("shape mismatch");
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *DenseMatrix<DT, MT>::copy2Device() {
    assert(!this->onDevice && this->data != nullptr);
    return new DenseMatrix<DT, MT>(this, true);
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *DenseMatrix<DT, MT>::copy2Host() {
    assert(this->onDevice && this->data != nullptr);
    return new DenseMatrix<DT, MT>(this, false);
}

template <typename DT, typename MT>
bool DenseMatrix<DT, MT>::toOrdering(ORDERING newOrdering) {
    if (this->ordering == newOrdering) return true;
    if (newOrdering != ORDERING::ROW_MAJOR && newOrdering != ORDERING::COL_MAJOR) throw std::runtime_error("Incorrect ordering value");
    // stored as [rows x cols] (row-major) or [cols x rows] (col-major): switching is a transpose
    const MT srcRows = this->ordering == ORDERING::ROW_MAJOR ? this->numRows : this->numCols;
    const MT srcCols = this->ordering == ORDERING::ROW_MAJOR ? this->numCols : this->numRows;
    DT *fresh = allocZeroed<DT>(numElements(), this->onDevice);
    if (this->onDevice) {
        if constexpr (std::is_same_v<DT, float>) {
            cuspmmCheck(cuspmm_transpose_f32(this->data, srcRows, srcCols, fresh, nullptr));
            cudaCheckError(cudaDeviceSynchronize());
        } else {
            throw std::runtime_error("device transpose is implemented for float only");
        }
    } else {
        for (MT r = 0; r < srcRows; ++r)
            for (MT c = 0; c < srcCols; ++c) fresh[(size_t)c * srcRows + r] = this->data[(size_t)r * srcCols + c];
    }
    freeSpaceOf(this->data, this->onDevice);
    this->data = fresh;
    this->ordering = newOrdering;
    return true;
}

template <typename DT, typename MT>
bool DenseMatrix<DT, MT>::save2File(std::string filePath) {
    DenseMatrix<DT, MT> *host = this->onDevice ? copy2Host() : nullptr;
    const DT *p = host ? host->data : this->data;
    std::ofstream out(filePath);
    if (!out.is_open()) {
        std::cerr << "Cannot open output file " << filePath << std::endl;
        delete host;
        return false;
    }
    if (this->ordering == ORDERING::ROW_MAJOR) {
        out << this->numRows << ' ' << this->numCols << std::endl;
        for (MT r = 0; r < this->numRows; ++r) {
            for (MT c = 0; c < this->numCols; ++c) out << p[RowMjIdx(r, c, this->numCols)] << ' ';
            out << std::endl;
        }
    } else {
        out << this->numRows << ' ' << this->numCols << " COL_MAJOR" << std::endl;
        for (MT c = 0; c < this->numCols; ++c) {
            for (MT r = 0; r < this->numRows; ++r) out << p[ColMjIdx(r, c, this->numRows)] << ' ';
            out << std::endl;
        }
    }
    delete host;
    return true;
}

template <typename DT, typename MT>
bool DenseMatrix<DT, MT>::allocateSpace(bool onDevice) {
    assert(this->data == nullptr);
    this->data = allocZeroed<DT>(numElements(), onDevice);
    this->onDevice = onDevice;
    return true;
}

template <typename DT, typename MT>
bool DenseMatrix<DT, MT>::freeSpace() {
    freeSpaceOf(this->data, this->onDevice);
    return true;
}

// =============================================================================== CSR
// *.csr: "rows cols nnz" / rowPtrs / colIdxs / values, one line each (src/formats/sparse_csr.cu:12-51)
template <typename DT, typename MT>
SparseMatrixCSR<DT, MT>::SparseMatrixCSR(std::string filePath) {
    std::ifstream in = openOrThrow(filePath);
    std::string line;
    in >> this->numRows >> this->numCols >> this->numNonZero;
    std::getline(in, line);
    allocateSpace(false);
    std::getline(in, line);
    { std::istringstream s(line); for (MT i = 0; i <= this->numRows; ++i) s >> rowPtrs[i]; }
    std::getline(in, line);
    { std::istringstream s(line); for (MT i = 0; i < this->numNonZero; ++i) s >> colIdxs[i]; }
    std::getline(in, line);
    { std::istringstream s(line); for (MT i = 0; i < this->numNonZero; ++i) s >> this->data[i]; }
}

template <typename DT, typename MT>
SparseMatrixCSR<DT, MT>::SparseMatrixCSR(MT numRows, MT numCols, MT numNonZero, bool onDevice) {
    this->numRows = numRows;
    this->numCols = numCols;
    this->numNonZero = numNonZero;
    allocateSpace(onDevice);
}

template <typename DT, typename MT>
SparseMatrixCSR<DT, MT>::~SparseMatrixCSR() {
    freeSpaceOf(rowPtrs, this->onDevice);
    freeSpaceOf(colIdxs, this->onDevice);
    freeSpaceOf(this->data, this->onDevice);
}

template <typename DT, typename MT>
void SparseMatrixCSR<DT, MT>::setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) {
    CHECK_CUSPARSE(cusparseCreateCsr(matDescP, this->numRows, this->numCols, this->numNonZero, rowPtrs, colIdxs, this->data,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, cudaTypeOf<DT>()));
}

template <typename DT, typename MT>
cusparseSpMMAlg_t SparseMatrixCSR<DT, MT>::getCusparseAlg() { return CUSPARSE_SPMM_CSR_ALG2; }

template <typename DT, typename MT>
SparseMatrixCSR<DT, MT> *SparseMatrixCSR<DT, MT>::copy2Device() {
    assert(!this->onDevice && this->data != nullptr);
    auto *d = new SparseMatrixCSR<DT, MT>(this->numRows, this->numCols, this->numNonZero, true);
    cudaCheckError(cudaMemcpy(d->rowPtrs, rowPtrs, ((size_t)this->numRows + 1) * sizeof(MT), cudaMemcpyHostToDevice));
    cudaCheckError(cudaMemcpy(d->colIdxs, colIdxs, (size_t)this->numNonZero * sizeof(MT), cudaMemcpyHostToDevice));
    cudaCheckError(cudaMemcpy(d->data, this->data, (size_t)this->numNonZero * sizeof(DT), cudaMemcpyHostToDevice));
    return d;
}

template <typename DT, typename MT>
SparseMatrixCSR<DT, MT> *SparseMatrixCSR<DT, MT>::copy2Host() {
    assert(this->onDevice);
    auto *h = new SparseMatrixCSR<DT, MT>(this->numRows, this->numCols, this->numNonZero, false);
    cudaCheckError(cudaMemcpy(h->rowPtrs, rowPtrs, ((size_t)this->numRows + 1) * sizeof(MT), cudaMemcpyDeviceToHost));
    cudaCheckError(cudaMemcpy(h->colIdxs, colIdxs, (size_t)this->numNonZero * sizeof(MT), cudaMemcpyDeviceToHost));
    cudaCheckError(cudaMemcpy(h->data, this->data, (size_t)this->numNonZero * sizeof(DT), cudaMemcpyDeviceToHost));
    return h;
}

template <typename DT, typename MT>
bool SparseMatrixCSR<DT, MT>::allocateSpace(bool onDevice) {
    assert(this->data == nullptr);
    this->data = allocZeroed<DT>(this->numNonZero, onDevice);
    rowPtrs = allocZeroed<MT>((size_t)this->numRows + 1, onDevice);
    colIdxs = allocZeroed<MT>(this->numNonZero, onDevice);
    this->onDevice = onDevice;
    return true;
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *SparseMatrixCSR<DT, MT>::toDense() {
    assert(!this->onDevice);
    auto *dm = new DenseMatrix<DT, MT>(this->numRows, this->numCols, false);
    for (MT r = 0; r < this->numRows; ++r)
        for (MT i = rowPtrs[r]; i < rowPtrs[r + 1]; ++i) dm->data[RowMjIdx(r, colIdxs[i], dm->numCols)] = this->data[i];
    return dm;
}

template <typename DT, typename MT>
std::ostream &operator<<(std::ostream &out, SparseMatrixCSR<DT, MT> &m) {
    out << m.numRows << ' ' << m.numCols << ' ' << m.numNonZero << std::endl;
    for (size_t i = 0; i <= m.numRows; ++i) out << m.rowPtrs[i] << ' ';
    out << std::endl;
    for (size_t i = 0; i < m.numNonZero; ++i) out << m.colIdxs[i] << ' ';
    out << std::endl;
    for (size_t i = 0; i < m.numNonZero; ++i) out << m.data[i] << ' ';
    out << std::endl;
    return out;
}

// =============================================================================== COO
// *.coo: "rows cols nnz" then nnz lines "row col value" (src/formats/sparse_coo.cu:13-38)
template <typename DT, typename MT>
SparseMatrixCOO<DT, MT>::SparseMatrixCOO(std::string filePath) {
    std::ifstream in = openOrThrow(filePath);
    in >> this->numRows >> this->numCols >> this->numNonZero;
    in.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    allocateSpace(false);
    for (size_t i = 0; i < this->numNonZero; ++i) in >> rowIdxs[i] >> colIdxs[i] >> this->data[i];
}

template <typename DT, typename MT>
SparseMatrixCOO<DT, MT>::SparseMatrixCOO(MT numRows, MT numCols, MT numNonZero, bool onDevice) {
    this->numRows = numRows;
    this->numCols = numCols;
    this->numNonZero = numNonZero;
    allocateSpace(onDevice);
}

template <typename DT, typename MT>
SparseMatrixCOO<DT, MT>::~SparseMatrixCOO() {
    freeSpaceOf(rowIdxs, this->onDevice);
    freeSpaceOf(colIdxs, this->onDevice);
    freeSpaceOf(this->data, this->onDevice);
}

template <typename DT, typename MT>
void SparseMatrixCOO<DT, MT>::setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) {
    CHECK_CUSPARSE(cusparseCreateCoo(matDescP, this->numRows, this->numCols, this->numNonZero, rowIdxs, colIdxs, this->data,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, cudaTypeOf<DT>()));
}

template <typename DT, typename MT>
cusparseSpMMAlg_t SparseMatrixCOO<DT, MT>::getCusparseAlg() { return CUSPARSE_SPMM_COO_ALG4; }

template <typename DT, typename MT>
SparseMatrixCOO<DT, MT> *SparseMatrixCOO<DT, MT>::copy2Device() {
    assert(!this->onDevice && this->data != nullptr);
    auto *d = new SparseMatrixCOO<DT, MT>(this->numRows, this->numCols, this->numNonZero, true);
    cudaCheckError(cudaMemcpy(d->rowIdxs, rowIdxs, (size_t)this->numNonZero * sizeof(MT), cudaMemcpyHostToDevice));
    cudaCheckError(cudaMemcpy(d->colIdxs, colIdxs, (size_t)this->numNonZero * sizeof(MT), cudaMemcpyHostToDevice));
    cudaCheckError(cudaMemcpy(d->data, this->data, (size_t)this->numNonZero * sizeof(DT), cudaMemcpyHostToDevice));
    return d;
}

template <typename DT, typename MT>
bool SparseMatrixCOO<DT, MT>::allocateSpace(bool onDevice) {
    assert(this->data == nullptr);
    this->data = allocZeroed<DT>(this->numNonZero, onDevice);
    rowIdxs = allocZeroed<MT>(this->numNonZero, onDevice);
    colIdxs = allocZeroed<MT>(this->numNonZero, onDevice);
    this->onDevice = onDevice;
    return true;
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *SparseMatrixCOO<DT, MT>::toDense() {
    assert(!this->onDevice);
    auto *dm = new DenseMatrix<DT, MT>(this->numRows, this->numCols, false);
    for (size_t i = 0; i < this->numNonZero; ++i) dm->data[RowMjIdx(rowIdxs[i], colIdxs[i], dm->numCols)] = this->data[i];
    return dm;
}

// =============================================================================== ELL (column-ELL)
// *_rowind.ell: "rows cols nnz max_nnz" then one line per COLUMN of max_nnz row indices (-1 = pad);
// *_values_colmajor.ell: the values, no header (src/formats/sparse_ell.cu:13-55)
template <typename DT, typename MT>
SparseMatrixELL<DT, MT>::SparseMatrixELL(std::string rowindPath, std::string valuesPath) {
    std::ifstream idx = openOrThrow(rowindPath);
    std::ifstream val = openOrThrow(valuesPath);
    idx >> this->numRows >> this->numCols >> this->numNonZero >> maxColNnz;
    idx.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    allocateSpace(false);
    const size_t slots = (size_t)this->numCols * maxColNnz;
    for (size_t i = 0; i < slots; ++i) {
        long long v = 0;           // "-1" must become 0xFFFFFFFF, as operator>> into uint32_t does
        idx >> v;
        rowIdxs[i] = (MT)v;
    }
    for (size_t i = 0; i < slots; ++i) val >> this->data[i];
}

template <typename DT, typename MT>
SparseMatrixELL<DT, MT>::SparseMatrixELL(MT numRows, MT numCols, MT numNonZero, MT maxColNnz, bool onDevice) {
    this->numRows = numRows;
    this->numCols = numCols;
    this->numNonZero = numNonZero;
    this->maxColNnz = maxColNnz;
    allocateSpace(onDevice);
}

template <typename DT, typename MT>
SparseMatrixELL<DT, MT>::~SparseMatrixELL() {
    freeSpaceOf(rowIdxs, this->onDevice);
    freeSpaceOf(this->data, this->onDevice);
}

template <typename DT, typename MT>
void SparseMatrixELL<DT, MT>::setCusparseSpMatDesc(cusparseSpMatDescr_t *) {
    throw std::runtime_error("not implemented");      // as in the reference (src/formats/sparse_ell.cu:105)
}

template <typename DT, typename MT>
cusparseSpMMAlg_t SparseMatrixELL<DT, MT>::getCusparseAlg() { return CUSPARSE_SPMM_ALG_DEFAULT; }

template <typename DT, typename MT>
bool SparseMatrixELL<DT, MT>::allocateSpace(bool onDevice) {
    assert(this->data == nullptr && rowIdxs == nullptr);
    const size_t slots = (size_t)this->numCols * maxColNnz;
    this->data = allocZeroed<DT>(slots, onDevice);
    rowIdxs = allocZeroed<MT>(slots, onDevice);
    this->onDevice = onDevice;
    return true;
}

template <typename DT, typename MT>
SparseMatrixELL<DT, MT> *SparseMatrixELL<DT, MT>::copy2Device() {
    assert(!this->onDevice && this->data != nullptr);
    auto *d = new SparseMatrixELL<DT, MT>(this->numRows, this->numCols, this->numNonZero, maxColNnz, true);
    const size_t slots = (size_t)this->numCols * maxColNnz;
    cudaCheckError(cudaMemcpy(d->rowIdxs, rowIdxs, slots * sizeof(MT), cudaMemcpyHostToDevice));
    cudaCheckError(cudaMemcpy(d->data, this->data, slots * sizeof(DT), cudaMemcpyHostToDevice));
    return d;
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *SparseMatrixELL<DT, MT>::toDense() {
    assert(!this->onDevice);
    auto *dm = new DenseMatrix<DT, MT>(this->numRows, this->numCols, false);
    for (size_t col = 0; col < this->numCols; ++col)
        for (size_t s = 0; s < maxColNnz; ++s) {
            const int row = (int)rowIdxs[col * maxColNnz + s];
            if (row >= 0) dm->data[RowMjIdx(row, col, dm->numCols)] = this->data[col * maxColNnz + s];
        }
    return dm;
}

template <typename DT, typename MT>
SlicedELL<DT, MT> *SparseMatrixELL<DT, MT>::toSliced() {
    if constexpr (!std::is_same_v<DT, float> || !std::is_same_v<MT, uint32_t>) {
        throw std::runtime_error("sliced ELL is implemented for <float, uint32_t>");
    } else {
        assert(this->onDevice);
        // column-ELL -> CSR (stable device radix sort by row) -> sliced ELL, all on the device
        SparseMatrixCSR<DT, MT> csr(this->numRows, this->numCols, this->numNonZero, true);
        cuspmmCheck(cuspmm_colell_to_csr(rowIdxs, this->data, this->numRows, this->numCols, maxColNnz, this->numNonZero,
                                         csr.rowPtrs, csr.colIdxs, csr.data, nullptr));
        auto *s = new SlicedELL<DT, MT>();
        s->numRows = this->numRows; s->numCols = this->numCols; s->numNonZero = this->numNonZero;
        s->numSlices = (this->numRows + 31) / 32;
        cudaCheckError(cudaMalloc(&s->slicePtrs, ((size_t)s->numSlices + 1) * sizeof(MT)));
        uint32_t slots = 0;
        cuspmmCheck(cuspmm_csr_to_sell_count(csr.rowPtrs, this->numRows, 32, s->slicePtrs, &slots, nullptr));
        s->numSlots = slots;
        cudaCheckError(cudaMalloc(&s->colIdxs, (size_t)(slots ? slots : 1) * sizeof(MT)));
        cudaCheckError(cudaMalloc(&s->data, (size_t)(slots ? slots : 1) * sizeof(DT)));
        cuspmmCheck(cuspmm_csr_to_sell_fill(csr.rowPtrs, csr.colIdxs, csr.data, this->numRows, 32, s->slicePtrs, s->colIdxs,
                                            s->data, nullptr));
        cudaCheckError(cudaDeviceSynchronize());
        return s;
    }
}

// =============================================================================== BSR
// *.bsr: "rows cols nnz br bc numBlocks" / block row ptrs / block col idx / block values
// (src/formats/sparse_bsr.cu:18-61)
template <typename DT, typename MT>
SparseMatrixBSR<DT, MT>::SparseMatrixBSR(std::string filePath) {
    std::ifstream in = openOrThrow(filePath);
    std::string line;
    in >> this->numRows >> this->numCols >> this->numNonZero >> blockRowSize >> blockColSize >> numBlocks;
    if (blockRowSize == 0 || blockColSize == 0) throw std::runtime_error("bad block size in " + filePath);
    numBlockRows = this->numRows / blockRowSize;
    numElements = numBlocks * blockRowSize * blockColSize;
    std::getline(in, line);
    allocateSpace(false);
    std::getline(in, line);
    { std::istringstream s(line); for (MT i = 0; i <= numBlockRows; ++i) s >> blockRowPtrs[i]; }
    std::getline(in, line);
    { std::istringstream s(line); for (MT i = 0; i < numBlocks; ++i) s >> blockColIdxs[i]; }
    for (MT i = 0; i < numElements; ++i) in >> this->data[i];
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT>::SparseMatrixBSR(MT numRows, MT numCols, MT numNonZero, MT blockRowSize, MT blockColSize,
                                         MT numBlocks, bool onDevice) {
    this->numRows = numRows;
    this->numCols = numCols;
    this->numNonZero = numNonZero;
    this->blockRowSize = blockRowSize;
    this->blockColSize = blockColSize;
    this->numBlocks = numBlocks;
    this->numBlockRows = blockRowSize ? numRows / blockRowSize : 0;   // the reference leaves this uninitialised (sparse_bsr.cu:78)
    this->numElements = numBlocks * blockRowSize * blockColSize;
    allocateSpace(onDevice);
    assertCheck();
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT>::SparseMatrixBSR(SparseMatrixBSR<DT, MT> *target, bool onDevice)
    : SparseMatrixBSR(target->numRows, target->numCols, target->numNonZero, target->blockRowSize, target->blockColSize,
                      target->numBlocks, onDevice) {
    copyData(target, onDevice);
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT>::~SparseMatrixBSR() {
    freeSpaceOf(blockRowPtrs, this->onDevice);
    freeSpaceOf(blockColIdxs, this->onDevice);
    freeSpaceOf(this->data, this->onDevice);
}

template <typename DT, typename MT>
void SparseMatrixBSR<DT, MT>::setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) {
    CHECK_CUSPARSE(cusparseCreateBsr(matDescP, numBlockRows, this->numCols / blockColSize, numBlocks, blockRowSize, blockColSize,
                                     blockRowPtrs, blockColIdxs, this->data, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_BASE_ZERO, cudaTypeOf<DT>(), CUSPARSE_ORDER_ROW));
}

template <typename DT, typename MT>
cusparseSpMMAlg_t SparseMatrixBSR<DT, MT>::getCusparseAlg() { return CUSPARSE_SPMM_ALG_DEFAULT; }

template <typename DT, typename MT>
bool SparseMatrixBSR<DT, MT>::copyData(SparseMatrixBSR<DT, MT> *source, bool onDevice) {
    assertSameShape(source);
    const cudaMemcpyKind k = kindOf(source->onDevice, onDevice);
    cudaCheckError(cudaMemcpy(blockRowPtrs, source->blockRowPtrs, ((size_t)numBlockRows + 1) * sizeof(MT), k));
    cudaCheckError(cudaMemcpy(blockColIdxs, source->blockColIdxs, (size_t)numBlocks * sizeof(MT), k));
    cudaCheckError(cudaMemcpy(this->data, source->data, (size_t)numElements * sizeof(DT), k));
    return true;
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT> *SparseMatrixBSR<DT, MT>::copy2Device() {
    assert(!this->onDevice && this->data != nullptr);
    return new SparseMatrixBSR<DT, MT>(this, true);
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT> *SparseMatrixBSR<DT, MT>::copy2Host() {
    assert(this->onDevice);
    return new SparseMatrixBSR<DT, MT>(this, false);
}

template <typename DT, typename MT>
void SparseMatrixBSR<DT, MT>::assertCheck() {
    if (blockRowSize == 0 || blockColSize == 0 || this->numRows % blockRowSize || this->numCols % blockColSize)
        throw std::runtime_error("BSR shape must be a multiple of the block shape");
}

template <typename DT, typename MT>
void SparseMatrixBSR<DT, MT>::assertSameShape(SparseMatrixBSR<DT, MT> *t) {
    if (!(blockRowSize == t->blockRowSize && blockColSize == t->blockColSize && numBlocks == t->numBlocks &&
          numBlockRows == t->numBlockRows && this->numRows == t->numRows && this->numCols == t->numCols))
        throw std::runtime_error("BSR shape mismatch");
}

template <typename DT, typename MT>
bool SparseMatrixBSR<DT, MT>::allocateSpace(bool onDevice) {
    assert(this->data == nullptr);
    this->data = allocZeroed<DT>(numElements, onDevice);
    blockRowPtrs = allocZeroed<MT>((size_t)numBlockRows + 1, onDevice);
    blockColIdxs = allocZeroed<MT>(numBlocks, onDevice);
    this->onDevice = onDevice;
    return true;
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT> *SparseMatrixBSR<DT, MT>::fromDense(DenseMatrix<DT, MT> *dense, MT br, MT bc) {
    if (dense->onDevice || dense->ordering != ORDERING::ROW_MAJOR || br == 0 || bc == 0 || dense->numRows % br || dense->numCols % bc)
        throw std::runtime_error("fromDense needs a row-major host matrix whose shape is a multiple of the block shape");
    const MT nbr = dense->numRows / br, nbc = dense->numCols / bc;
    std::vector<MT> ptr(nbr + 1, 0), cols;
    for (MT R = 0; R < nbr; ++R) {
        for (MT Cb = 0; Cb < nbc; ++Cb) {
            bool any = false;
            for (MT i = 0; i < br && !any; ++i)
                for (MT j = 0; j < bc && !any; ++j) any = dense->data[RowMjIdx(R * br + i, Cb * bc + j, dense->numCols)] != DT(0);
            if (any) cols.push_back(Cb);
        }
        ptr[R + 1] = (MT)cols.size();
    }
    auto *m = new SparseMatrixBSR<DT, MT>(dense->numRows, dense->numCols, (MT)(cols.size() * br * bc), br, bc, (MT)cols.size(), false);
    std::copy(ptr.begin(), ptr.end(), m->blockRowPtrs);
    std::copy(cols.begin(), cols.end(), m->blockColIdxs);
    for (MT R = 0; R < nbr; ++R)
        for (MT b = ptr[R]; b < ptr[R + 1]; ++b)
            for (MT i = 0; i < br; ++i)
                for (MT j = 0; j < bc; ++j)
                    m->data[((size_t)b * br + i) * bc + j] = dense->data[RowMjIdx(R * br + i, cols[b] * bc + j, dense->numCols)];
    return m;
}

template <typename DT, typename MT>
SparseMatrixBSR<DT, MT> *SparseMatrixBSR<DT, MT>::fromCSR(SparseMatrixCSR<DT, MT> *csr, MT br, MT bc) {
    if constexpr (!std::is_same_v<DT, float> || !std::is_same_v<MT, uint32_t>) {
        throw std::runtime_error("device CSR->BSR is implemented for <float, uint32_t>");
    } else {
        assert(csr->onDevice);
        const MT Mp = (csr->numRows + br - 1) / br * br, Kp = (csr->numCols + bc - 1) / bc * bc;
        MT *ptr = nullptr;
        cudaCheckError(cudaMalloc(&ptr, ((size_t)Mp / br + 1) * sizeof(MT)));
        uint32_t nb = 0;
        cuspmmCheck(cuspmm_csr_to_bsr_count(csr->rowPtrs, csr->colIdxs, csr->numRows, csr->numCols, csr->numNonZero, br, bc, ptr,
                                            &nb, nullptr));
        auto *m = new SparseMatrixBSR<DT, MT>(Mp, Kp, nb * br * bc, br, bc, nb, true);
        cudaCheckError(cudaMemcpy(m->blockRowPtrs, ptr, ((size_t)Mp / br + 1) * sizeof(MT), cudaMemcpyDeviceToDevice));
        cudaCheckError(cudaFree(ptr));
        cuspmmCheck(cuspmm_csr_to_bsr_fill(csr->rowPtrs, csr->colIdxs, csr->data, csr->numRows, csr->numCols, csr->numNonZero, br,
                                           bc, nb, m->blockColIdxs, m->data, nullptr));
        cudaCheckError(cudaDeviceSynchronize());
        return m;
    }
}

template <typename DT, typename MT>
DenseMatrix<DT, MT> *SparseMatrixBSR<DT, MT>::toDense() {
    assert(!this->onDevice);
    auto *dm = new DenseMatrix<DT, MT>(this->numRows, this->numCols, false);
    for (MT R = 0; R < numBlockRows; ++R)
        for (MT b = blockRowPtrs[R]; b < blockRowPtrs[R + 1]; ++b)
            for (MT i = 0; i < blockRowSize; ++i)
                for (MT j = 0; j < blockColSize; ++j)   // block stride = blockColSize, the layout spmmBSRCpu reads
                    dm->data[RowMjIdx(R * blockRowSize + i, blockColIdxs[b] * blockColSize + j, dm->numCols)] =
                        this->data[((size_t)b * blockRowSize + i) * blockColSize + j];
    return dm;
}

template class DenseMatrix<float, uint32_t>;
template class DenseMatrix<double, uint32_t>;
template class SparseMatrixCSR<float, uint32_t>;
template class SparseMatrixCSR<double, uint32_t>;
template class SparseMatrixCOO<float, uint32_t>;
template class SparseMatrixCOO<double, uint32_t>;
template class SparseMatrixELL<float, uint32_t>;
template class SparseMatrixELL<double, uint32_t>;
template class SparseMatrixBSR<float, uint32_t>;
template class SparseMatrixBSR<double, uint32_t>;
template std::ostream &operator<<(std::ostream &, SparseMatrixCSR<float, uint32_t> &);

}  // namespace cuspmm
