/*
 * qt_shim_core.h — stand-ins for the Qt 5 / Qt3D / OpenCL names the reference's CPU path mentions.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Purpose: compile the reference's OWN, UNMODIFIED sources
 *   src/CCPUParticleSimulator.cpp, src/CBaseParticleSimulator.cpp, src/CCollisionGeometry.cpp,
 *   src/CGrid.cpp, src/CParticle.cpp, src/renderableentity.cpp, src/CScene.cpp
 * where they lie under /root/reference into oracle/_ref/libsph_ref.so (oracle/Makefile, target `ref`),
 * so that oracle/sph_oracle.cpp (the hand restatement) can be required to match the reference
 * bit for bit.  Nothing under gmu-water-simulation_b200/ includes this directory.
 *
 * What is real and what is restated:
 *   - every line of simulation code (loops, narrowing, traversal order, swap-and-pop, wall bounce,
 *     scene generators) is the reference's, compiled as is;
 *   - the render-side classes (QEntity, QTransform, QSphereMesh, materials ...) are empty shells: the
 *     reference only constructs and parents them;
 *   - QVector3D — the ONE piece of arithmetic that lives in Qt — is restated below from Qt 5's
 *     qtbase/src/gui/math3d/qvector3d.h / qvector3d.cpp, function by function (Qt is not in this image
 *     and there is no network, so the bodies are quoted from the published Qt 5 sources by function name,
 *     not by line; CMakeLists.txt:60 pins no version, these bodies are the same throughout Qt 5.5–5.15);
 *   - Qt3DExtras::QCuboidGeometry yields the vertex/index buffers Qt's createCuboidVertexData()
 *     produces for the default 2x2 face resolution: 24 vertices (position 3f, texcoord 2f, normal 3f,
 *     tangent 4f; stride 48 B) at exactly +-extent/2, 36 ushort indices — so the reference's own
 *     CCollisionGeometry::init() extracts the bounding-box walls from it, unmodified.
 *
 * Only float overloads of the QVector3D operators exist, exactly as in Qt: a double operand in the
 * reference's expressions is narrowed to float by the compiler's implicit conversion, not by us.
 */
#ifndef SPH_QT_SHIM_CORE_H
#define SPH_QT_SHIM_CORE_H

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- moc / QObject vocabulary ------------------------------------------------------------- */
#define Q_OBJECT
#define signals public
#define slots
#define emit
#define SIGNAL(a) "2" #a
#define SLOT(a) "1" #a

typedef long long qint64;
typedef unsigned int QRgb;

namespace Qt {
enum Key { Key_Space = 0x20, Key_G = 0x47, Key_O = 0x4f, Key_P = 0x50, Key_R = 0x52, Key_S = 0x53 };
enum Initialization { Uninitialized };
}

class QObject
{
public:
    /* parent/child ownership as in Qt: a parent deletes its children (qobject.cpp, QObjectPrivate::deleteChildren) */
    explicit QObject(QObject *parent = nullptr) : m_parent(nullptr) { setParent(parent); }
    virtual ~QObject()
    {
        std::vector<QObject *> kids;
        kids.swap(m_children);
        for (QObject *k : kids) { k->m_parent = nullptr; delete k; }
        if (m_parent) {
            std::vector<QObject *> &v = m_parent->m_children;
            std::vector<QObject *>::iterator it = std::find(v.begin(), v.end(), this);
            if (it != v.end()) { *it = v.back(); v.pop_back(); }
        }
    }
    QObject *parent() const { return m_parent; }
    void setParent(QObject *p)
    {
        if (m_parent == p) return;
        if (m_parent) {
            std::vector<QObject *> &v = m_parent->m_children;
            std::vector<QObject *>::iterator it = std::find(v.begin(), v.end(), this);
            if (it != v.end()) { *it = v.back(); v.pop_back(); }
        }
        m_parent = p;
        if (p) p->m_children.push_back(this);
    }
    /* headless: the 0-ms QTimer never fires, the driver calls step() itself */
    static bool connect(const QObject *, const char *, const QObject *, const char *) { return true; }
private:
    QObject(const QObject &);
    QObject &operator=(const QObject &);
    QObject *m_parent;
    std::vector<QObject *> m_children;
};

class QString
{
public:
    QString() {}
    QString(const char *s) : m_s(s ? s : "") {}
    QString(const std::string &s) : m_s(s) {}
    const std::string &toStdString() const { return m_s; }
    bool operator==(const QString &o) const { return m_s == o.m_s; }
    bool operator!=(const QString &o) const { return m_s != o.m_s; }
private:
    std::string m_s;
};

class QColor
{
public:
    QColor() {}
    QColor(QRgb) {}
    QColor(const char *) {}
};

template<typename A, typename B> using QPair = std::pair<A, B>;

class QByteArray
{
public:
    QByteArray() {}
    QByteArray(const char *p, int n) : m_d(p, p + n) {}
    char *data() { return m_d.data(); }
    const char *data() const { return m_d.data(); }
    int size() const { return (int) m_d.size(); }
private:
    std::vector<char> m_d;
};

template<typename T>
class QVector
{
public:
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    void resize(int n) { m_v.resize((size_t) n); }
    void clear() { m_v.clear(); }
    void push_back(const T &t) { m_v.push_back(t); }
    int size() const { return (int) m_v.size(); }
    T &operator[](int i) { return m_v[(size_t) i]; }
    const T &operator[](int i) const { return m_v[(size_t) i]; }
    const T &at(int i) const { return m_v.at((size_t) i); }
    iterator begin() { return m_v.begin(); }
    iterator end() { return m_v.end(); }
    const_iterator begin() const { return m_v.begin(); }
    const_iterator end() const { return m_v.end(); }
private:
    std::vector<T> m_v;
};

template<typename T>
class QList
{
public:
    QList &operator<<(const T &t) { m_v.push_back(t); return *this; }
    int size() const { return (int) m_v.size(); }
    const T &at(int i) const { return m_v.at((size_t) i); }
    typename std::vector<T>::const_iterator begin() const { return m_v.begin(); }
    typename std::vector<T>::const_iterator end() const { return m_v.end(); }
    void clear() { m_v.clear(); }
private:
    std::vector<T> m_v;
};

class QDebug
{
public:
    QDebug &nospace() { return *this; }
    template<typename T> QDebug &operator<<(const T &) { return *this; }
};
class QDebugStateSaver
{
public:
    explicit QDebugStateSaver(QDebug &) {}
};
inline QDebug qDebug() { return QDebug(); }

class QTimer : public QObject
{
public:
    void start() { m_active = true; }
    void stop() { m_active = false; }
    bool isActive() const { return m_active; }
private:
    bool m_active = false;
};

class QElapsedTimer
{
public:
    void start() { m_t0 = std::chrono::steady_clock::now(); }
    qint64 restart() { qint64 e = elapsed(); start(); return e; }
    /* integer milliseconds, like Qt */
    qint64 elapsed() const
    {
        return (qint64) std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - m_t0).count();
    }
private:
    std::chrono::steady_clock::time_point m_t0 = std::chrono::steady_clock::now();
};

inline unsigned int qNextPowerOfTwo(unsigned int v)
{
    /* qmath.h: strictly greater power of two */
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}

/* ---- QVector3D: restated from Qt 5 (qtbase/src/gui/math3d) ----------------------------------
 * Components are three floats xp, yp, zp (private members of QVector3D in qvector3d.h).  The inline operators below
 * are the bodies of the `inline` definitions at the end of qvector3d.h (operator+=, -=, *=, /=, the friend operators
 * +, -, *, unary -, /); length / lengthSquared / dotProduct / normalize are the out-of-line bodies of qvector3d.cpp
 * (QVector3D::length(), ::lengthSquared(), ::dotProduct(), ::normalize()).  No operator takes a double. */
class QVector3D
{
public:
    QVector3D() : xp(0.0f), yp(0.0f), zp(0.0f) {}
    QVector3D(float xpos, float ypos, float zpos) : xp(xpos), yp(ypos), zp(zpos) {}

    float x() const { return xp; }
    float y() const { return yp; }
    float z() const { return zp; }
    void setX(float aX) { xp = aX; }
    void setY(float aY) { yp = aY; }
    void setZ(float aZ) { zp = aZ; }

    /* qvector3d.cpp: "Need some extra precision if the length is very small." — fp64 sum, one narrowing */
    float length() const
    {
        double len = double(xp) * double(xp) + double(yp) * double(yp) + double(zp) * double(zp);
        return float(std::sqrt(len));
    }
    /* qvector3d.cpp: return xp * xp + yp * yp + zp * zp;  (fp32, left to right) */
    float lengthSquared() const { return xp * xp + yp * yp + zp * zp; }

    /* qvector3d.cpp normalize(): fp64 squared length; unchanged if qFuzzyIsNull(len - 1.0f) or
     * qFuzzyIsNull(len) (qglobal.h: |d| <= 0.000000000001); else each component divided in fp64 by sqrt(len) */
    void normalize()
    {
        double len = double(xp) * double(xp) + double(yp) * double(yp) + double(zp) * double(zp);
        if (std::fabs(len - 1.0f) <= 0.000000000001 || std::fabs(len) <= 0.000000000001)
            return;
        len = std::sqrt(len);
        xp = float(double(xp) / len);
        yp = float(double(yp) / len);
        zp = float(double(zp) / len);
    }

    QVector3D &operator+=(const QVector3D &vector) { xp += vector.xp; yp += vector.yp; zp += vector.zp; return *this; }
    QVector3D &operator-=(const QVector3D &vector) { xp -= vector.xp; yp -= vector.yp; zp -= vector.zp; return *this; }
    QVector3D &operator*=(float factor) { xp *= factor; yp *= factor; zp *= factor; return *this; }
    QVector3D &operator/=(float divisor) { xp /= divisor; yp /= divisor; zp /= divisor; return *this; }

    /* qvector3d.cpp: return v1.xp * v2.xp + v1.yp * v2.yp + v1.zp * v2.zp; */
    static float dotProduct(const QVector3D &v1, const QVector3D &v2) { return v1.xp * v2.xp + v1.yp * v2.yp + v1.zp * v2.zp; }

    friend inline const QVector3D operator+(const QVector3D &v1, const QVector3D &v2) { return QVector3D(v1.xp + v2.xp, v1.yp + v2.yp, v1.zp + v2.zp); }
    friend inline const QVector3D operator-(const QVector3D &v1, const QVector3D &v2) { return QVector3D(v1.xp - v2.xp, v1.yp - v2.yp, v1.zp - v2.zp); }
    friend inline const QVector3D operator*(float factor, const QVector3D &vector) { return QVector3D(vector.xp * factor, vector.yp * factor, vector.zp * factor); }
    friend inline const QVector3D operator*(const QVector3D &vector, float factor) { return QVector3D(vector.xp * factor, vector.yp * factor, vector.zp * factor); }
    friend inline const QVector3D operator*(const QVector3D &v1, const QVector3D &v2) { return QVector3D(v1.xp * v2.xp, v1.yp * v2.yp, v1.zp * v2.zp); }
    friend inline const QVector3D operator-(const QVector3D &vector) { return QVector3D(-vector.xp, -vector.yp, -vector.zp); }
    friend inline const QVector3D operator/(const QVector3D &vector, float divisor) { return QVector3D(vector.xp / divisor, vector.yp / divisor, vector.zp / divisor); }
    friend inline const QVector3D operator/(const QVector3D &vector, const QVector3D &divisor) { return QVector3D(vector.xp / divisor.xp, vector.yp / divisor.yp, vector.zp / divisor.zp); }

private:
    float xp, yp, zp;
};

/* ---- OpenCL scalar / vector typedefs (CL/cl_platform.h): cl_float3 is cl_float4, 16-byte aligned ---- */
typedef float cl_float;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef cl_uint cl_bool;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_device_type;
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)

typedef union
{
    cl_float __attribute__((aligned(16))) s[4];
    __extension__ struct { cl_float x, y, z, w; };
} cl_float4;
typedef cl_float4 cl_float3;

typedef union
{
    cl_int __attribute__((aligned(16))) s[4];
    __extension__ struct { cl_int x, y, z, w; };
} cl_int4;
typedef cl_int4 cl_int3;

namespace cl {
/* only named as member / parameter types by CLWrapper.h, CLPlatforms.h, CGPUBaseParticleSimulator.h */
class Device {};
class Context {};
class CommandQueue {};
class Program {};
class Platform {};
class Event {};
class Kernel {};
class Buffer {};
class NDRange {};
static const NDRange NullRange;
}

/* ---- Qt3D shells -------------------------------------------------------------------------- */
namespace Qt3DCore {
class QNode : public QObject
{
public:
    explicit QNode(QNode *parent = nullptr) : QObject(parent) {}
    void setEnabled(bool) {}
};
class QComponent : public QNode
{
public:
    explicit QComponent(QNode *parent = nullptr) : QNode(parent) {}
};
class QEntity : public QNode
{
public:
    explicit QEntity(QNode *parent = nullptr) : QNode(parent) {}
    /* qentity.cpp addComponent(): a parentless component is adopted by the entity */
    void addComponent(QComponent *c);
};
class QTransform : public QComponent
{
public:
    explicit QTransform(QNode *parent = nullptr) : QComponent(parent) {}
    void setTranslation(const QVector3D &t) { m_translation = t; }
    void setScale(float) {}
    QVector3D translation() const { return m_translation; }
private:
    QVector3D m_translation;
};
inline void QEntity::addComponent(QComponent *c) { if (c && !c->parent()) c->setParent(this); }
}

namespace Qt3DRender {

class QBufferDataGenerator
{
public:
    virtual ~QBufferDataGenerator() {}
    virtual QByteArray operator()() = 0;
};
typedef std::shared_ptr<QBufferDataGenerator> QBufferDataGeneratorPtr;

class QBuffer : public Qt3DCore::QNode
{
public:
    explicit QBuffer(Qt3DCore::QNode *parent = nullptr) : Qt3DCore::QNode(parent) {}
    void setDataGenerator(const QBufferDataGeneratorPtr &g) { m_gen = g; }
    QBufferDataGeneratorPtr dataGenerator() const { return m_gen; }
private:
    QBufferDataGeneratorPtr m_gen;
};

class QAttribute : public Qt3DCore::QNode
{
public:
    enum AttributeType { VertexAttribute, IndexAttribute };
    enum VertexBaseType { Byte = 0, UnsignedByte, Short, UnsignedShort, Int, UnsignedInt, HalfFloat, Float, Double };

    QAttribute(QBuffer *buf, const QString &name, VertexBaseType vbt, unsigned vertexSize, unsigned count,
               unsigned offset, unsigned stride, AttributeType type)
        : m_buffer(buf), m_name(name), m_vbt(vbt), m_vertexSize(vertexSize), m_count(count), m_offset(offset),
          m_stride(stride), m_type(type) {}

    static QString defaultPositionAttributeName() { return QString("vertexPosition"); }
    static QString defaultNormalAttributeName() { return QString("vertexNormal"); }
    QString name() const { return m_name; }
    AttributeType attributeType() const { return m_type; }
    VertexBaseType vertexBaseType() const { return m_vbt; }
    unsigned vertexSize() const { return m_vertexSize; }
    unsigned count() const { return m_count; }
    unsigned byteOffset() const { return m_offset; }
    unsigned byteStride() const { return m_stride; }
    QBuffer *buffer() const { return m_buffer; }
private:
    QBuffer *m_buffer;
    QString m_name;
    VertexBaseType m_vbt;
    unsigned m_vertexSize, m_count, m_offset, m_stride;
    AttributeType m_type;
};

class QGeometry : public Qt3DCore::QNode
{
public:
    explicit QGeometry(Qt3DCore::QNode *parent = nullptr) : Qt3DCore::QNode(parent) {}
    QVector<QAttribute *> attributes() const { return m_attributes; }
    void addAttribute(QAttribute *a) { m_attributes.push_back(a); }
private:
    QVector<QAttribute *> m_attributes;
};

class QGeometryRenderer : public Qt3DCore::QComponent
{
public:
    enum PrimitiveType { Points = 0, Lines = 1, Triangles = 4 };
    explicit QGeometryRenderer(Qt3DCore::QNode *parent = nullptr) : Qt3DCore::QComponent(parent) {}
    void setGeometry(QGeometry *) {}
    void setPrimitiveType(PrimitiveType) {}
};
class QMesh : public QGeometryRenderer
{
public:
    explicit QMesh(Qt3DCore::QNode *parent = nullptr) : QGeometryRenderer(parent) {}
};
class QMaterial : public Qt3DCore::QComponent
{
public:
    explicit QMaterial(Qt3DCore::QNode *parent = nullptr) : Qt3DCore::QComponent(parent) {}
};
class QEffect;
class QTechnique;
class QRenderPass;
class QShaderProgram;
class QPointLight : public Qt3DCore::QComponent
{
public:
    explicit QPointLight(Qt3DCore::QNode *parent = nullptr) : Qt3DCore::QComponent(parent) {}
    void setColor(const QColor &) {}
    void setIntensity(float) {}
};
}

namespace Qt3DExtras {
class QSphereMesh : public Qt3DRender::QGeometryRenderer
{
public:
    explicit QSphereMesh(Qt3DCore::QNode *parent = nullptr) : Qt3DRender::QGeometryRenderer(parent) {}
    void setRings(int) {}
    void setSlices(int) {}
    void setRadius(float) {}
};
class QPhongMaterial : public Qt3DRender::QMaterial
{
public:
    explicit QPhongMaterial(Qt3DCore::QNode *parent = nullptr) : Qt3DRender::QMaterial(parent) {}
    void setDiffuse(const QColor &) {}
};

/* Qt3DExtras::QCuboidGeometry (qt3d/src/extras/geometries/qcuboidgeometry.cpp): with the default face
 * resolution QSize(2, 2) createCuboidVertexData() emits, for each of the six planes (+x, -x, +y, -y, +z, -z),
 * the 2x2 corner vertices a = a0 + i*da, b = b0 + j*db with a0 = -extent/2.0f, da = extent/(res-1), so every
 * coordinate is exactly -e/2 or fl(-e/2 + e) = +e/2 in fp32.  Element = position(3) texcoord(2) normal(3)
 * tangent(4) floats, stride 48 bytes; indices are 36 unsigned shorts (two triangles per face). */
class QCuboidGeometry : public Qt3DRender::QGeometry
{
public:
    explicit QCuboidGeometry(Qt3DCore::QNode *parent = nullptr) : Qt3DRender::QGeometry(parent), m_x(1.0f), m_y(1.0f), m_z(1.0f)
    {
        m_vertexBuffer = new Qt3DRender::QBuffer(this);
        m_indexBuffer = new Qt3DRender::QBuffer(this);
        const unsigned stride = (3 + 2 + 3 + 4) * sizeof(float);
        typedef Qt3DRender::QAttribute A;
        addAttribute(new A(m_vertexBuffer, A::defaultPositionAttributeName(), A::Float, 3, 24, 0, stride, A::VertexAttribute));
        addAttribute(new A(m_vertexBuffer, QString("vertexTexCoord"), A::Float, 2, 24, 3 * sizeof(float), stride, A::VertexAttribute));
        addAttribute(new A(m_vertexBuffer, A::defaultNormalAttributeName(), A::Float, 3, 24, 5 * sizeof(float), stride, A::VertexAttribute));
        addAttribute(new A(m_vertexBuffer, QString("vertexTangent"), A::Float, 4, 24, 8 * sizeof(float), stride, A::VertexAttribute));
        addAttribute(new A(m_indexBuffer, QString(), A::UnsignedShort, 1, 36, 0, 0, A::IndexAttribute));
        updateVertices();
        updateIndices();
    }
    void setXExtent(float v) { m_x = v; }
    void setYExtent(float v) { m_y = v; }
    void setZExtent(float v) { m_z = v; }
    void updateVertices();
    void updateIndices();
private:
    float m_x, m_y, m_z;
    Qt3DRender::QBuffer *m_vertexBuffer, *m_indexBuffer;
};

namespace shim_detail {
struct BytesGenerator : Qt3DRender::QBufferDataGenerator
{
    QByteArray bytes;
    explicit BytesGenerator(const QByteArray &b) : bytes(b) {}
    QByteArray operator()() override { return bytes; }
};
/* one plane of createCuboidVertexData(): in-plane extents (w along a, h along b), plane position pc on the normal axis */
inline void plane(std::vector<float> &out, int axisA, int axisB, int axisN, float w, float h, float pc, float sign)
{
    const float a0 = -w / 2.0f, b0 = -h / 2.0f, da = w / 1.0f, db = h / 1.0f;
    for (int j = 0; j < 2; ++j) {
        const float b = b0 + static_cast<float>(j) * db;
        for (int i = 0; i < 2; ++i) {
            const float a = a0 + static_cast<float>(i) * da;
            float p[3], n[3] = {0.0f, 0.0f, 0.0f};
            p[axisA] = a; p[axisB] = b; p[axisN] = pc;
            n[axisN] = sign;
            out.insert(out.end(), {p[0], p[1], p[2]});
            out.insert(out.end(), {static_cast<float>(i), static_cast<float>(j)});
            out.insert(out.end(), {n[0], n[1], n[2]});
            out.insert(out.end(), {0.0f, 0.0f, 0.0f, 1.0f});
        }
    }
}
}

inline void QCuboidGeometry::updateVertices()
{
    std::vector<float> v;
    v.reserve(24 * 12);
    shim_detail::plane(v, 2, 1, 0, m_z, m_y, m_x / 2.0f, 1.0f);    /* +x : yz plane */
    shim_detail::plane(v, 2, 1, 0, m_z, m_y, -m_x / 2.0f, -1.0f);  /* -x */
    shim_detail::plane(v, 0, 2, 1, m_x, m_z, m_y / 2.0f, 1.0f);    /* +y : xz plane */
    shim_detail::plane(v, 0, 2, 1, m_x, m_z, -m_y / 2.0f, -1.0f);  /* -y */
    shim_detail::plane(v, 0, 1, 2, m_x, m_y, m_z / 2.0f, 1.0f);    /* +z : xy plane */
    shim_detail::plane(v, 0, 1, 2, m_x, m_y, -m_z / 2.0f, -1.0f);  /* -z */
    QByteArray bytes(reinterpret_cast<const char *>(v.data()), (int) (v.size() * sizeof(float)));
    m_vertexBuffer->setDataGenerator(Qt3DRender::QBufferDataGeneratorPtr(new shim_detail::BytesGenerator(bytes)));
}

inline void QCuboidGeometry::updateIndices()
{
    std::vector<unsigned short> idx;
    for (unsigned short f = 0; f < 6; ++f) {
        const unsigned short b = (unsigned short) (4 * f);
        const unsigned short tri[6] = {b, (unsigned short) (b + 1), (unsigned short) (b + 3), b, (unsigned short) (b + 3), (unsigned short) (b + 2)};
        idx.insert(idx.end(), tri, tri + 6);
    }
    QByteArray bytes(reinterpret_cast<const char *>(idx.data()), (int) (idx.size() * sizeof(unsigned short)));
    m_indexBuffer->setDataGenerator(Qt3DRender::QBufferDataGeneratorPtr(new shim_detail::BytesGenerator(bytes)));
}
}

#endif /* SPH_QT_SHIM_CORE_H */
