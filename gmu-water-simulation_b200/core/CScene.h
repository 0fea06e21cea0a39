// CScene.h — headless stand-in for the Qt3D scene (include/CScene.h:19).
// The simulator constructors take a CScene*; without the viewer there is no entity tree, so the root
// entity is null.  The optional Qt3D viewer supplies the real class instead of this one.
#pragma once

namespace Qt3DCore { class QEntity; }

class CScene {
public:
    CScene() = default;
    Qt3DCore::QEntity *getRootEntity() const { return nullptr; }
};
